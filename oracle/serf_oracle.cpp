// serf_oracle.cpp — CPU restatement of serf-core's membership state-merge path.
//
//   *** TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT. ***
//   Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
//   legs may load this library.  The product (serf_b200/csrc) never links or calls it.
//
// Parity status
//   * serf half (LamportClock, intent buffer, handle_node_{join,leave}[_intent], push-pull
//     ordering, reaper, queue cap): PINNED by the reference's own known-answer tests,
//     replayed literally in tests/test_oracle_kat.py (list: SURVEY.md §8c).
//   * SWIM half (alive/suspect/dead merge, incarnation, suspicion timer, retransmit limit,
//     gossip fan-out): the algorithm lives in the un-vendored dependency
//     memberlist-core = "0.8.1" (reference Cargo.toml:40, no Cargo.lock), a port of
//     hashicorp/memberlist.  Restated here from the published SWIM + Lifeguard algorithm
//     and upstream behaviour; the reference tree holds no test that pins it:
//     PARITY UNPINNED for that half.
//
// Two models live here:
//   Part A  RefNode   — ONE serf node, literal: HashMap-like tables keyed by id, the
//                       recent-intent buffer, left/failed lists, a TransmitLimitedQueue.
//                       This is what the reference KATs are replayed against.
//   Part B  TickSim   — N virtual nodes × R tracked subjects, bulk-synchronous ticks, the
//                       message set of every destination applied ONE MESSAGE AT A TIME in
//                       the canonical order (DESIGN.md §"Tick semantics").  The CUDA path
//                       must reproduce its records, clocks and per-tick trace bit-exactly.
//
// All `file:line` citations are relative to /root/reference/serf-core/src/.
//
// Build: see oracle/Makefile (g++ -O2 -shared -fPIC).  Plain C ABI at the bottom (ctypes).

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <pthread.h>
#include <sched.h>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/serfsim.h"   // config / stats / trace-row struct layouts and constants only

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// =====================================================================================
// Shared primitives
// =====================================================================================

// MemberStatus — types/member.rs:54-58
enum : u8 { ST_NONE = 0, ST_ALIVE = 1, ST_LEAVING = 2, ST_LEFT = 3, ST_FAILED = 4 };
// MessageType tags — types/message.rs:17-18
enum : u8 { TY_NONE = 0, TY_LEAVE = 1, TY_JOIN = 2 };
// memberlist node states (external crate; restated)
enum : u8 { ML_ALIVE = 0, ML_SUSPECT = 1, ML_DEAD = 2, ML_LEFT = 3 };
// SerfState — serf.rs (Alive, Leaving, Left, Shutdown)
enum : u8 { SS_ALIVE = 0, SS_LEAVING = 1, SS_LEFT = 2 };

// LamportClock — types/clock.rs:121-173
struct LamportClock {
  u64 v = 0;
  u64 time() const { return v; }                        // clock.rs:142-145
  u64 increment() { v += 1; return v; }                 // clock.rs:148-151 (fetch_add(1) + 1)
  void witness(u64 t) {                                 // clock.rs:155-172
    if (t < v) return;                                  // "If the other value is old, we do not need to do anything"
    v = t + 1;                                          // CAS(cur → t + 1)
  }
};

// memberlist retransmit limit: retransmit_mult * ceil(log10(n + 1))   [external; SURVEY §8c]
static u32 retransmit_limit(u32 mult, u64 n) {
  u32 digits = 0;          // smallest k with 10^k >= n + 1  == ceil(log10(n+1)) in exact arithmetic
  u64 p = 1;
  while (p < n + 1) { p *= 10; ++digits; }
  return mult * digits;
}

// memberlist suspicion timeouts in ticks [external; SURVEY §8c]:
//   min = suspicion_mult * max(1, log10(max(1,n))) * probe_interval   (ms arithmetic as upstream:
//         mult * int(node_scale*1000) * interval / 1000),  max = suspicion_max_timeout_mult * min,
//   k   = suspicion_mult - 2 (0 when n - 2 < k),
//   timeout(c) = max - log(c+1)/log(k+1) * (max - min), floored to ms, >= min;  k < 1 → min.
// Returned table has k+1 entries, each ceil(ms / tick_ms), at least 1.
static std::vector<u32> suspicion_table(u32 susp_mult, u32 max_mult, u32 probe_ticks, u32 tick_ms, u64 n) {
  double node_scale = std::max(1.0, std::log10(std::max(1.0, (double)n)));
  int64_t interval_ms = (int64_t)probe_ticks * tick_ms;
  int64_t min_ms = (int64_t)susp_mult * (int64_t)(node_scale * 1000.0) * interval_ms / 1000;
  int64_t max_ms = (int64_t)max_mult * min_ms;
  int64_t k = (int64_t)susp_mult - 2;
  if ((int64_t)n - 2 < k) k = 0;
  if (k < 0) k = 0;
  std::vector<u32> tab;
  for (int64_t c = 0; c <= k; ++c) {
    int64_t ms;
    if (k < 1) ms = min_ms;
    else {
      double frac = std::log((double)c + 1.0) / std::log((double)k + 1.0);
      double raw = (double)max_ms - frac * (double)(max_ms - min_ms);
      ms = (int64_t)std::floor(raw);
      if (ms < min_ms) ms = min_ms;
    }
    int64_t ticks = (ms + tick_ms - 1) / tick_ms;
    if (ticks < 1) ticks = 1;
    tab.push_back((u32)ticks);
  }
  return tab;
}

// Philox4x32-10 (Salmon et al., SC'11), the counter RNG both sides use.
static inline void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 out[4]) {
  const u32 M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    u64 p0 = (u64)M0 * c0, p1 = (u64)M1 * c2;
    u32 hi0 = (u32)(p0 >> 32), lo0 = (u32)p0, hi1 = (u32)(p1 >> 32), lo1 = (u32)p1;
    u32 n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }

enum : u32 { DOMAIN_GOSSIP = 0, DOMAIN_PROBE = 1, DOMAIN_PUSHPULL = 2 };

static inline u64 mix64(u64 x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31; return x;
}
static inline u32 from_hash(u32 node) { return (node * 0x9E3779B1u) >> 28; }   // 4-bit confirmer bucket

// =====================================================================================
// Part A — RefNode: one serf node, literal tables
// =====================================================================================

struct MemberStateA { u8 status; u64 status_time; bool has_leave_time; int64_t leave_time_ms; };   // types/member.rs:20-26
struct NodeIntentA { u8 ty; int64_t wall_ms; u64 ltime; };                                          // types/member.rs:30-34
struct QueuedA { u8 ty; u64 ltime; u64 id; bool prune; u32 transmits; u64 seq; u32 len; bool notify; };

struct RefNode {
  u64 self_id;
  u8 serf_state = SS_ALIVE;
  LamportClock clock, event_clock, query_clock;                 // serf.rs:138-140
  std::map<u64, MemberStateA> states;                           // Members.states          types/member.rs:37
  std::map<u64, NodeIntentA> recent_intents;                    // Members.recent_intents  types/member.rs:38
  std::vector<std::pair<u64, MemberStateA>> left_members, failed_members;   // types/member.rs:39-40
  std::vector<QueuedA> broadcasts;                              // intent TransmitLimitedQueue, base.rs:179-190
  u32 retransmit_mult = 4;
  u64 seq = 0;
  u32 refutes = 0;          // number of spawned broadcast_join refutations (base.rs:1470-1480)
  std::vector<std::pair<u32, u64>> events;   // (MemberEventType, id) in emission order
  // user-event ring ("next" row 3; restated for the KATs only): buffer[ltime % len] = {ltime, [(name, payload)…]}
  struct UserEvents { u64 ltime; std::vector<std::pair<std::string, std::string>> events; };
  std::vector<std::pair<bool, UserEvents>> event_buffer = std::vector<std::pair<bool, UserEvents>>(512);   // options.rs:517 event_buffer_size
  u64 event_min_time = 0;
  std::vector<std::pair<std::string, std::string>> user_events_out;          // delivered to the application, in order

  explicit RefNode(u64 id) : self_id(id) {
    // base.rs:198-200: all three clocks are incremented once at construction.
    clock.increment(); event_clock.increment(); query_clock.increment();
    // base.rs:269-272: notify_join(local) inserts self (Alive, status_time 0).
    states[id] = MemberStateA{ST_ALIVE, 0, false, 0};
  }

  // base.rs:1817-1866 -----------------------------------------------------------------
  bool upsert_intent(u64 id, u8 ty, u64 ltime, int64_t wall_ms) {
    auto it = recent_intents.find(id);
    if (it != recent_intents.end()) {                       // Entry::Occupied
      if (ltime > it->second.ltime) {
        it->second.ty = ty; it->second.ltime = ltime; it->second.wall_ms = wall_ms;
        return true;
      }
      return false;
    }
    recent_intents[id] = NodeIntentA{ty, wall_ms, ltime};   // Entry::Vacant
    return true;
  }
  bool recent_intent(u64 id, u8 ty, u64* ltime) const {
    auto it = recent_intents.find(id);
    if (it != recent_intents.end() && it->second.ty == ty) { *ltime = it->second.ltime; return true; }
    return false;
  }
  void reap_intents(int64_t now_ms, int64_t timeout_ms) {   // retain (now - wall) <= timeout
    for (auto it = recent_intents.begin(); it != recent_intents.end();) {
      if ((now_ms - it->second.wall_ms) <= timeout_ms) ++it; else it = recent_intents.erase(it);
    }
  }
  static void remove_old_member(std::vector<std::pair<u64, MemberStateA>>& old, u64 id) {   // base.rs:1813-1815
    old.erase(std::remove_if(old.begin(), old.end(), [&](auto& m) { return m.first == id; }), old.end());
  }

  // base.rs:364-375 + the external TransmitLimitedQueue.queue_broadcast (SerfBroadcast never
  // invalidates: broadcast.rs:24-30).
  void queue(u8 ty, u64 ltime, u64 id, bool prune, bool notify) {
    broadcasts.push_back(QueuedA{ty, ltime, id, prune, 0, seq++, 24, notify});
  }

  // base.rs:1338-1373 -----------------------------------------------------------------
  bool handle_node_join_intent(u64 ltime, u64 id, int64_t now_ms = 0) {
    clock.witness(ltime);                                   // :1340
    auto it = states.find(id);
    if (it != states.end()) {
      MemberStateA& m = it->second;
      if (ltime <= m.status_time) return false;             // :1346
      m.status_time = ltime;                                // :1351
      if (m.status == ST_LEAVING) m.status = ST_ALIVE;      // :1356-1358
      return true;
    }
    return upsert_intent(id, TY_JOIN, ltime, now_ms);       // :1362-1370
  }

  // base.rs:381-397
  void broadcast_join(u64 ltime) {
    clock.witness(ltime);
    handle_node_join_intent(ltime, self_id);
    queue(TY_JOIN, ltime, self_id, false, false);
  }

  // base.rs:1442-1572 -----------------------------------------------------------------
  // Returns rebroadcast.  A refutation (":1470-1480", spawn_detach(broadcast_join)) is
  // recorded in `refute_pending`; run_detached() executes it (the reference runs it on
  // another task "since we have the memberLock").
  bool refute_pending = false;
  bool handle_node_leave_intent(u64 ltime, u64 id, bool prune, int64_t now_ms = 0) {
    u8 state = serf_state;                                  // :1443
    clock.witness(ltime);                                   // :1446
    auto it = states.find(id);
    if (it == states.end())                                 // :1450-1458
      return upsert_intent(id, TY_LEAVE, ltime, now_ms);
    MemberStateA& m = it->second;
    if (ltime <= m.status_time) return false;               // :1464
    if (id == self_id && state == SS_ALIVE) {               // :1470-1480 refute
      refute_pending = true; ++refutes;
      return false;
    }
    m.status_time = ltime;                                  // :1497 — always, before the switch
    switch (m.status) {
      case ST_NONE: return false;                           // :1501
      case ST_ALIVE:                                        // :1502-1511
        m.status = ST_LEAVING;
        if (prune) handle_prune(id);
        return true;
      case ST_LEAVING: case ST_LEFT:                        // :1512-1519
        if (prune) handle_prune(id);
        return true;
      case ST_FAILED: {                                     // :1520-1559
        m.status = ST_LEFT;
        MemberStateA owned = m;
        remove_old_member(failed_members, id);
        left_members.push_back({id, owned});
        events.push_back({SERFSIM_EVENT_LEAVE, id});
        if (prune) handle_prune(id);
        return true;
      }
      default:                                              // :1560-1570 unknown status
        m.status = ST_LEAVING;
        if (prune) handle_prune(id);
        return true;
    }
  }
  void run_detached() { if (refute_pending) { refute_pending = false; broadcast_join(clock.time()); } }

  // base.rs:1628-1653 (the Leaving-state sleep is wall-clock only; erase is immediate here)
  void handle_prune(u64 id) {
    auto it = states.find(id);
    if (it == states.end()) return;
    u8 ms = it->second.status;
    if (ms == ST_LEAVING || ms == ST_LEFT) remove_old_member(left_members, id);
    states.erase(it);                                       // erase_node! :499-519
    events.push_back({SERFSIM_EVENT_REAP, id});
  }

  // base.rs:1206-1334 -----------------------------------------------------------------
  void handle_node_join(u64 id) {
    auto it = states.find(id);
    u8 old_status;
    if (it != states.end()) {
      old_status = it->second.status;
      it->second.status = ST_ALIVE;                         // :1251-1263 (status_time kept, leave_time None)
      it->second.has_leave_time = false;
    } else {
      u8 status = ST_ALIVE; u64 status_ltime = 0, t;        // :1276-1288
      if (recent_intent(id, TY_JOIN, &t)) status_ltime = t;
      if (recent_intent(id, TY_LEAVE, &t)) { status_ltime = t; status = ST_LEAVING; }
      states[id] = MemberStateA{status, status_ltime, false, 0};
      old_status = ST_NONE;
    }
    events.push_back({SERFSIM_EVENT_JOIN, id});
    if (old_status == ST_FAILED || old_status == ST_LEFT) { // :1317-1320
      remove_old_member(failed_members, id);
      remove_old_member(left_members, id);
    }
  }

  // base.rs:1375-1440 -----------------------------------------------------------------
  void handle_node_leave(u64 id, int64_t now_ms) {
    auto it = states.find(id);
    if (it == states.end()) return;                         // :1378-1380
    MemberStateA& m = it->second;
    if (m.status == ST_LEAVING) {                           // :1384-1393
      m.status = ST_LEFT; m.has_leave_time = true; m.leave_time_ms = now_ms;
      left_members.push_back({id, m});
      events.push_back({SERFSIM_EVENT_LEAVE, id});
    } else if (m.status == ST_ALIVE) {                      // :1394-1402
      m.status = ST_FAILED; m.has_leave_time = true; m.leave_time_ms = now_ms;
      failed_members.push_back({id, m});
      events.push_back({SERFSIM_EVENT_FAILED, id});
    }                                                       // :1403-1406 anything else: warn + return
  }

  // serf/delegate.rs:427-554 (membership part) -----------------------------------------
  void merge_remote_state(u64 pp_ltime, const u64* ids, const u64* ltimes, u32 n, const u64* left, u32 n_left,
                          u64 event_ltime, u64 query_ltime) {
    if (pp_ltime > 0) clock.witness(pp_ltime - 1);          // :466-468
    if (event_ltime > 0) event_clock.witness(event_ltime - 1);
    if (query_ltime > 0) query_clock.witness(query_ltime - 1);
    auto find = [&](u64 id, u64* lt) { for (u32 i = 0; i < n; ++i) if (ids[i] == id) { *lt = ltimes[i]; return true; } return false; };
    for (u32 i = 0; i < n_left; ++i) {                      // :495-510 left nodes first, ltime + 1
      u64 lt;
      if (find(left[i], &lt)) { handle_node_leave_intent(lt + 1, left[i], false); }
    }
    for (u32 i = 0; i < n; ++i) {                           // :513-523 then artificial joins
      bool is_left = false;
      for (u32 j = 0; j < n_left; ++j) if (left[j] == ids[i]) is_left = true;
      if (is_left) continue;
      handle_node_join_intent(ltimes[i], ids[i]);
    }
  }

  // base.rs:750-837 handle_user_event → rebroadcast?  (quirk kept: an occupied ring slot is reused without checking
  // or refreshing its ltime, base.rs:801-813)
  bool handle_user_event(u64 ltime, const std::string& name, const std::string& payload) {
    event_clock.witness(ltime);                                       // :763
    if (ltime < event_min_time) return false;                         // :768-770
    const u64 bltime = event_buffer.size();
    const u64 cur = event_clock.time();
    if (cur > bltime && ltime < cur - bltime) return false;           // :773-783 too old
    auto& slot = event_buffer[(size_t)(ltime % bltime)];
    if (slot.first) {
      for (auto& prev : slot.second.events) if (prev.first == name && prev.second == payload) return false;   // :803-808 already seen
      slot.second.events.push_back({name, payload});
    } else {
      slot.first = true; slot.second.ltime = ltime; slot.second.events = {{name, payload}};
    }
    user_events_out.push_back({name, payload});
    return true;
  }

  // serf/delegate.rs:386-425 local_state: {ltime: clock.time(), status_ltimes: every member's status_time, left_members: ids of
  // the left list, event_ltime, query_ltime} (the event ring itself is out of scope)
  u32 local_state(u64* pp_ltime, u64* ids, u64* ltimes, u32 cap, u64* left, u32 cap_left, u32* n_left, u64* event_ltime, u64* query_ltime) const {
    *pp_ltime = clock.time(); *event_ltime = event_clock.time(); *query_ltime = query_clock.time();
    u32 n = 0;
    for (auto& kv : states) { if (n < cap) { ids[n] = kv.first; ltimes[n] = kv.second.status_time; } ++n; }
    u32 nl = 0;
    for (auto& m : left_members) { if (nl < cap_left) left[nl] = m.first; ++nl; }
    *n_left = nl;
    return n;
  }

  // Local API — serf/api.rs:318-361 (join, after memberlist.join), :422-499 (leave),
  // base.rs:454-480 (force_leave).  has_alive_members(): base.rs:346-359.
  bool has_alive_members() const {
    for (auto& kv : states) if (kv.first != self_id && kv.second.status == ST_ALIVE) return true;
    return false;
  }
  void api_join() { broadcast_join(clock.time()); }
  int api_leave() {
    if (serf_state == SS_LEFT) return 0;
    if (serf_state == SS_LEAVING) return -1;
    serf_state = SS_LEAVING;
    u64 t = clock.time();
    clock.increment();
    handle_node_leave_intent(t, self_id, false);
    if (has_alive_members()) queue(TY_LEAVE, t, self_id, false, true);
    return 0;
  }
  void api_force_leave(u64 id, bool prune) {
    u64 t = clock.time();
    handle_node_leave_intent(t, id, prune);
    if (!has_alive_members()) return;
    queue(TY_LEAVE, t, id, prune, true);
  }

  // External TransmitLimitedQueue.get_broadcasts restated: lowest transmit count first
  // (ties: longer message, then newer), each pick increments transmits, entries reaching
  // retransmit_limit(mult, NumMembers = states.len()) are dropped (finished()).
  // serf.rs:109-131 (NumMembers), serf/delegate.rs:317-344.
  u32 get_broadcasts(u32 byte_limit, u32 overhead, u8* ty_out, u64* lt_out, u64* id_out, u32 cap, u8* prune_out = nullptr) {
    u32 limit = retransmit_limit(retransmit_mult, states.size());
    std::vector<size_t> order(broadcasts.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      const QueuedA &x = broadcasts[a], &y = broadcasts[b];
      if (x.transmits != y.transmits) return x.transmits < y.transmits;
      if (x.len != y.len) return x.len > y.len;
      return x.seq > y.seq;
    });
    u32 used = 0, n = 0;
    std::vector<size_t> done;
    for (size_t i : order) {
      QueuedA& q = broadcasts[i];
      if (used + overhead + q.len > byte_limit) continue;
      used += overhead + q.len;
      if (n < cap) { ty_out[n] = q.ty; lt_out[n] = q.ltime; id_out[n] = q.id; if (prune_out) prune_out[n] = q.prune ? 1 : 0; }
      ++n;
      if (++q.transmits >= limit) done.push_back(i);
    }
    std::sort(done.begin(), done.end());
    for (size_t j = done.size(); j-- > 0;) broadcasts.erase(broadcasts.begin() + done[j]);
    return n;
  }

  // Reaper — base.rs:483-610: reap!(failed, reconnect_timeout), reap!(left, tombstone_timeout),
  // reap_intents(recent_intent_timeout).
  void reap_list(std::vector<std::pair<u64, MemberStateA>>& list, int64_t now_ms, int64_t timeout_ms) {
    size_t n = list.size(), i = 0;
    while (i < n) {
      auto m = list[i];
      if (m.second.has_leave_time && (now_ms - m.second.leave_time_ms) <= timeout_ms) { ++i; continue; }   // :536-541
      list[i] = list[n - 1]; list.pop_back(); --n;                                                        // swap_remove :544
      states.erase(m.first);                                                                              // erase_node!
      events.push_back({SERFSIM_EVENT_REAP, m.first});
    }
  }
  void reap(int64_t now_ms, int64_t reconnect_timeout_ms, int64_t tombstone_timeout_ms, int64_t intent_timeout_ms) {
    reap_list(failed_members, now_ms, reconnect_timeout_ms);
    reap_list(left_members, now_ms, tombstone_timeout_ms);
    reap_intents(now_ms, intent_timeout_ms);
  }
};

// QueueChecker::get_queue_max — base.rs:728-739
static u64 get_queue_max(u64 max_queue_depth, u64 min_queue_depth, u64 num_members) {
  u64 max = max_queue_depth;
  if (min_queue_depth > 0) {
    max = num_members * 2;
    if (max < min_queue_depth) max = min_queue_depth;
  }
  return max;
}

// =====================================================================================
// Part B — TickSim: N nodes × R subject views, literal per-message application
// =====================================================================================

#pragma pack(push, 1)
struct View {              // the 32-byte member record (DESIGN.md "Record layout")
  u32 st;                  //  0 status_time, or buffered-intent ltime while !known
  u32 qjoin;               //  4 ltime of the queued join intent
  u32 qleave;              //  8 ltime of the queued leave intent
  u32 inc;                 // 12 memberlist incarnation of the subject as seen
  u32 deadline;            // 16 tick at which the suspicion timer fires (0 = none)
  u32 leave_tick;          // 20 MemberState.leave_time as tick+1 (0 = None)
  u8 status;               // 24 MemberStatus, or buffered-intent type while !known
  u8 ml;                   // 25 bits 0-1 memberlist state, bits 2-5 from-hash of the queued suspect
  u8 txj, txl, txm;        // 26-28 remaining transmits: join intent, leave intent, memberlist message
  u8 flags;                // 29 bit0 = known (member present in Members.states)
  u16 mask;                // 30 suspicion confirmer buckets
};
#pragma pack(pop)
static_assert(sizeof(View) == 32, "record must be 32 bytes");

static inline u8 ml_state(const View& r) { return r.ml & 3; }
static inline void set_ml(View& r, u8 state, u8 fromh) { r.ml = (u8)((state & 3) | ((fromh & 15) << 2)); }
static inline bool known(const View& r) { return r.flags & 1; }

struct NodeB { u32 clock; u8 up; u8 sstate; };

struct RuleCtx {            // per-run constants
  u32 limit;                // retransmit limit
  u32 k;                    // max confirmations
  const u32* timeout;       // [k+1]
};

static inline void witness32(u32& c, u32 t) { if (t < c) return; c = t + 1; }   // clock.rs:155-172 on the device-width clock

// --- serf-layer rules on a view (same statements as RefNode, on the packed record) ------
static bool v_join_intent(View& r, u32 lt, const RuleCtx& cx, bool requeue = true) {            // base.rs:1338-1373 (witness done by caller)
  bool acc;
  if (known(r)) {
    if (lt <= r.st) return false;
    r.st = lt;
    if (r.status == ST_LEAVING) r.status = ST_ALIVE;
    acc = true;
  } else {                                                                   // upsert_intent, base.rs:1838-1866
    if (r.status == TY_NONE || lt > r.st) { r.status = TY_JOIN; r.st = lt; acc = true; } else acc = false;
  }
  if (acc && requeue) { r.qjoin = lt; r.txj = (u8)cx.limit; }                // serf/delegate.rs:294-300 re-queue (push-pull discards the result, :495-523)
  return acc;
}
static bool v_leave_intent(View& r, u32 lt, bool self, u8 sstate, bool* refute, const RuleCtx& cx, bool requeue = true, bool prune = false) {   // base.rs:1442-1572
  bool acc;
  if (!known(r)) {
    if (r.status == TY_NONE || lt > r.st) { r.status = TY_LEAVE; r.st = lt; acc = true; } else acc = false;
  } else {
    if (lt <= r.st) return false;
    if (self && sstate == SS_ALIVE) { *refute = true; return false; }
    r.st = lt;
    switch (r.status) {
      case ST_NONE: acc = false; break;
      case ST_ALIVE: r.status = ST_LEAVING; acc = true; break;
      case ST_LEAVING: case ST_LEFT: acc = true; break;
      case ST_FAILED: r.status = ST_LEFT; acc = true; break;
      default: r.status = ST_LEAVING; acc = true; break;
    }
    // msg.prune → handle_prune (base.rs:1504-1570, 1628-1653): the member is erased from the table (erase_node!, :499-519).  The
    // reference first sleeps broadcast_timeout + leave_propagate_delay when the member is Leaving (holding the member lock); the
    // tick model erases at once, like RefNode::handle_prune above.
    if (acc && prune) { r.flags &= (u8)~1u; r.status = TY_NONE; r.st = 0; r.leave_tick = 0; }
  }
  if (acc && requeue) { r.qleave = lt; r.txl = (u8)cx.limit; r.flags = (u8)((r.flags & ~2u) | (prune ? 2u : 0u)); }   // flags bit 1: the queued leave intent carries prune
  return acc;
}
static void v_node_join(View& r) {                                           // base.rs:1206-1334
  if (known(r)) { r.status = ST_ALIVE; r.leave_tick = 0; return; }
  u8 status = ST_ALIVE; u32 st = 0;
  if (r.status == TY_JOIN) st = r.st;
  if (r.status == TY_LEAVE) { st = r.st; status = ST_LEAVING; }
  r.status = status; r.st = st; r.flags |= 1; r.leave_tick = 0;
}
static void v_node_leave(View& r, u32 tick) {                                // base.rs:1375-1440
  if (!known(r)) return;
  if (r.status == ST_LEAVING) { r.status = ST_LEFT; r.leave_tick = tick + 1; }
  else if (r.status == ST_ALIVE) { r.status = ST_FAILED; r.leave_tick = tick + 1; }
}

// --- memberlist rules on a view (external crate; restated from hashicorp/memberlist state.go
//     aliveNode / suspectNode / deadNode / refute, which memberlist-core ports) -----------
static void v_refute(View& r, u32 accused, const RuleCtx& cx) {
  u32 inc = r.inc + 1;
  if (accused >= inc) inc = accused + 1;
  r.inc = inc;
  set_ml(r, ML_ALIVE, 0);
  r.txm = (u8)cx.limit;
}
static void v_ml_alive(View& r, u32 a, bool self, const RuleCtx& cx) {
  if (a <= r.inc) return;                       // non-local: old incarnation; local: a < inc ignored, a == inc "same version"
  if (self) { v_refute(r, a, cx); return; }
  r.deadline = 0; r.mask = 0;                   // delete(nodeTimers)
  u8 old = ml_state(r);
  r.inc = a; set_ml(r, ML_ALIVE, 0); r.txm = (u8)cx.limit;
  if (old == ML_DEAD || old == ML_LEFT) v_node_join(r);     // EventDelegate::notify_join, serf/delegate.rs:565
}
static void v_ml_suspect(View& r, u32 s, u32 fromh, u32 tick, bool self, const RuleCtx& cx) {
  if (s < r.inc) return;
  if (ml_state(r) == ML_SUSPECT) {              // timer exists → Confirm(from)
    u32 n_old = (u32)__builtin_popcount(r.mask) - 1;
    if (n_old >= cx.k) return;
    if (r.mask & (1u << fromh)) return;
    r.mask |= (u16)(1u << fromh);
    r.deadline = r.deadline - cx.timeout[n_old] + cx.timeout[n_old + 1];
    set_ml(r, ML_SUSPECT, (u8)fromh); r.txm = (u8)cx.limit;
    return;
  }
  if (ml_state(r) != ML_ALIVE) return;
  if (self) { v_refute(r, s, cx); return; }
  r.inc = s; set_ml(r, ML_SUSPECT, (u8)fromh); r.mask = (u16)(1u << fromh);
  r.deadline = tick + cx.timeout[0]; r.txm = (u8)cx.limit;
}
static void v_ml_dead(View& r, u32 d, bool left, u32 tick, bool self, const RuleCtx& cx) {
  if (d < r.inc) return;
  r.deadline = 0; r.mask = 0;                   // delete(nodeTimers)
  u8 s = ml_state(r);
  if (s == ML_DEAD || s == ML_LEFT) return;
  if (self) { v_refute(r, d, cx); return; }     // a node that has left is already ML_LEFT (returned above)
  r.inc = d; set_ml(r, left ? ML_LEFT : ML_DEAD, 0); r.txm = (u8)cx.limit;
  v_node_leave(r, tick);                        // EventDelegate::notify_leave, serf/delegate.rs:571
}
static inline u32 ml_key(const View& r) { return (r.inc << 6) | ((u32)ml_state(r) << 4) | ((r.ml >> 2) & 15); }

// byzantine injector rules (BASELINE configs[4]; defined here, there is no reference behaviour to follow)
static void byz_stale(const View& r, u32 delta, u8* kind, u32* lt, u32* key) {        // the aged copy of a view an injector re-sends
  *lt = r.st > delta ? r.st - delta : 0;
  const u32 inc = r.inc > delta ? r.inc - delta : 0;
  *kind = (r.status == ST_LEAVING || r.status == ST_LEFT) ? 0 : 1;
  *key = (inc << 6) | ((u32)ml_state(r) << 4) | ((r.ml >> 2) & 15);
}
static bool byz_judge(const View& q, u8 kind, u32 val, u32 delta) {                    // receiver's verdict on ONE arriving stale entry
  return (kind == 2) ? (q.inc >= (val >> 6) + delta) : (known(q) && q.st >= val + delta);
}

struct Msg { u32 dst, src, val; u8 slot, kind, byz = 0, prune = 0; };   // prune: LeaveMessage.prune (types/leave.rs:39-44), leave intents only   // kind 0 leave, 1 join, 2 memberlist; byz: a stale entry injected by a byzantine node

// ---- user events: literal per-node state ----
// EventCore.buffer is `Vec<Option<UserEvents>>` of event_buffer_size = 512 entries (serf/base.rs:193, options.rs:516);
// kept here as a map ring-index → slot, which is the same thing without the 512 × N empty entries.
struct UeSlotB { u32 ltime; std::vector<u32> events; };      // UserEvents { ltime, events } — events hold tracked-event indices
struct UeNodeB {
  u32 clock = 1;                                             // event_clock after Serf::new (serf/base.rs:198-200)
  std::map<u32, UeSlotB> ring;
  u8 tx[8] = {0, 0, 0, 0, 0, 0, 0, 0};                       // remaining transmits of the queued broadcast of tracked event e
};
struct UeMsg { u32 dst, e; };
struct EventB { u32 tick, op, node, slot; };

// Persistent workers of the threaded tick loop (test infrastructure: the timed CPU baseline and the full-size checks).  A tick used to create and
// join one thread per worker — twice with anti-entropy on —, i.e. 64 thread creations and affinity calls per tick on the bench hosts: for a study
// whose ticks are mostly idle that was most of the run time, and it made the timings depend on the box (round-1 review: "not a stable anchor").
// The workers are now created once per instance, pinned once and parked on a condition variable between ticks.
struct WorkerPool {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  const std::function<void(u32)>* job = nullptr;
  unsigned long long gen = 0;
  u32 left = 0;
  bool stop = false;
  WorkerPool(u32 T, const std::vector<int>& pin) {
    for (u32 c = 0; c < T; ++c) th.emplace_back([this, c, pin] {
      if (!pin.empty()) { cpu_set_t set; CPU_ZERO(&set); CPU_SET(pin[c % pin.size()], &set); pthread_setaffinity_np(pthread_self(), sizeof(set), &set); }
      unsigned long long seen = 0;
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv_go.wait(lk, [&] { return stop || gen != seen; });
        if (stop) return;
        seen = gen;
        const std::function<void(u32)>* f = job;
        lk.unlock();
        (*f)(c);
        lk.lock();
        if (--left == 0) cv_done.notify_one();
      }
    });
  }
  void run(const std::function<void(u32)>& f) {              // f(c) on every worker c; returns when all are done
    std::unique_lock<std::mutex> lk(m);
    job = &f; left = (u32)th.size(); ++gen;
    cv_go.notify_all();
    cv_done.wait(lk, [&] { return left == 0; });
  }
  ~WorkerPool() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv_go.notify_all();
    for (auto& t : th) t.join();
  }
};

struct TickSim {
  serfsim_config_t cfg;
  u32 N, R, tick = 0;
  std::vector<u64> row_ptr; std::vector<u32> col;
  std::vector<u32> subj;
  std::vector<View> rec;        // [R][N]
  std::vector<NodeB> node;
  std::vector<EventB> events;   // sorted by (tick, insertion)
  std::unordered_set<u64> event_keys;
  std::unordered_map<u32, std::vector<EventB>> ev_by_tick;
  u32 max_event_tick = 0; bool any_event = false;
  std::vector<u32> timeout; RuleCtx cx;
  std::vector<serfsim_tick_row_t> trace;
  std::vector<u8> subj_up;      // ground truth per slot
  std::vector<u16> watch;       // per node: bit s set iff subject s is in the node's neighbour list (only those nodes can probe it)
  std::vector<std::vector<Msg>> mail;   // messages sent in the previous tick: [producer range][consumer range]
  std::vector<std::vector<Msg>> mail_next;                       // the boxes being filled this tick (capacity is reused)
  u32 own_first = 0, own_count = 0;                              // id range this instance owns (sharded runs; default: everything)
  std::vector<std::vector<Msg>> exported;                        // per thread: messages of the last tick for nodes owned elsewhere
  struct Scratch { std::vector<u32> head, pos; std::vector<Msg> byd; };
  std::vector<Scratch> scratch;                                  // per-thread buffers, reused across ticks
  u32 chunk = 1;
  u64 tot_events = 0;
  // user events
  u32 ue_n = 0; u32 ue_content[8] = {0}; u32 ue_ltime[8] = {0}; u32 ue_injected = 0;
  u32 ue_stamped = 0; bool ue_alias_error = false;             // events whose origin has fired; two of them in one ring slot with different ltimes
  std::vector<UeNodeB> uen;
  std::vector<std::vector<UeMsg>> ue_mail, ue_mail_next;       // [producer range][consumer range], like `mail`
  u64 ue_tot[5] = {0, 0, 0, 0, 0};                             // messages, edge_updates, delivered, duplicates, too_old
  // byzantine stale-record injectors (BASELINE configs[4]; no reference semantics — this block IS the definition)
  std::vector<u8> byz;                                         // per node: 1 = injector
  u32 byz_n = 0, byz_delta = 2;
  std::vector<u8> anomaly;                                     // per node: sender flag
  u64 byz_tot[3] = {0, 0, 0};                                  // injected entries, injected (peer, subject) pairs, senders flagged
  int threads = 1;
  std::unique_ptr<WorkerPool> pool; std::vector<int> pool_pin;   // workers of the threaded tick loop (created on first use, re-created when threads / pin change)
  void run_workers(u32 T, const std::function<void(u32)>& f) {
    if (!pool || pool->th.size() != T || pool_pin != pin) { pool.reset(); pool.reset(new WorkerPool(T, pin)); pool_pin = pin; }
    pool->run(f);
  }
  std::vector<int> pin;                                        // optional: worker c of the tick loop runs on logical CPU pin[c % size] (stable timings)
  std::string err;
  void pin_worker(u32 c) const {
    if (pin.empty()) return;
    cpu_set_t set; CPU_ZERO(&set); CPU_SET(pin[c % pin.size()], &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }

  View& at(u32 s, u32 v) { return rec[(size_t)s * N + v]; }

  void init_tables() {
    timeout = suspicion_table(cfg.suspicion_mult, cfg.suspicion_max_timeout_mult,
                              cfg.probe_interval_ticks ? cfg.probe_interval_ticks : 1, cfg.gossip_interval_ms, N);
    cx.limit = retransmit_limit(cfg.retransmit_mult, N);
    cx.k = (u32)timeout.size() - 1;
    cx.timeout = timeout.data();
  }
  void reset(u64 seed) {
    cfg.seed = seed; tick = 0; events.clear(); event_keys.clear(); ev_by_tick.clear(); any_event = false; max_event_tick = 0; mail_next.clear(); trace.clear(); mail.clear(); tot_events = 0;
    chunk = (N + threads - 1) / threads;
    if (own_count == 0) { own_first = 0; own_count = N; }
    rec.assign((size_t)R * N, View{});
    node.assign(N, NodeB{cfg.init_clock, 1, SS_ALIVE});
    for (u32 s = 0; s < R; ++s)
      for (u32 v = 0; v < N; ++v) {
        View& r = at(s, v);
        r.st = cfg.init_status_ltime; r.inc = 1; r.status = ST_ALIVE; r.flags = 1; set_ml(r, ML_ALIVE, 0);
      }
    subj_up.assign(R, 1);
    ue_reset();
    anomaly.assign(byz_n ? N : 0, 0); for (auto& x : byz_tot) x = 0;
  }
  void ue_reset() {
    uen.assign(ue_n ? N : 0, UeNodeB{}); ue_mail.clear(); ue_mail_next.clear(); ue_injected = 0; ue_stamped = 0; ue_alias_error = false;
    for (auto& x : ue_ltime) x = 0;
    for (auto& x : ue_tot) x = 0;
  }
  // Serf::handle_user_event (serf/base.rs:750-837) for tracked event e carrying Lamport time L.
  // 0 accepted (→ rebroadcast), 1 duplicate, 2 too old.
  int ue_handle(UeNodeB& nd, u32 e, u32 L) {
    witness32(nd.clock, L);                                                    // :763
    const u32 min_time = 0;                                                    // EventCore.min_time: only moved by a join with event_join_ignore (delegate.rs:531-537)
    if (L < min_time) return 2;                                                // :766-768
    const u32 bltime = 512, cur = nd.clock;                                    // :771-772
    if (cur > bltime && L < cur - bltime) return 2;                            // :773-781
    const u32 idx = L % bltime;                                                // :784
    auto it = nd.ring.find(idx);
    if (it != nd.ring.end()) {                                                 // occupied: the slot's own ltime is not looked at (quirk ii)
      for (u32 prev : it->second.events) if (ue_content[prev] == ue_content[e]) return 1;   // :801-806 user_event.eq(prev): name and payload
      it->second.events.push_back(e);                                          // :807
    } else {
      nd.ring[idx] = UeSlotB{L, {e}};                                          // :809-813
    }
    return 0;                                                                  // :836 true
  }
  void ue_record(u32 v, u32 out[4]) const {                    // the 16-byte event record the CUDA path keeps
    const UeNodeB& nd = uen[v];
    u32 seen = 0, first = 0;
    // `first` is a derived field: the lowest-index event of each occupied ring slot — independent of arrival order
    for (auto& kv : nd.ring) { u32 lo = 0xffffffffu; for (u32 e : kv.second.events) { seen |= 1u << e; lo = std::min(lo, e); } first |= 1u << lo; }
    out[0] = nd.clock; out[1] = seen | (first << 8); out[2] = 0; out[3] = 0;
    for (u32 e = 0; e < 4; ++e) { out[2] |= (u32)nd.tx[e] << (8 * e); out[3] |= (u32)nd.tx[4 + e] << (8 * e); }
  }
  void compute_watch() {
    watch.assign(N, 0);
    if (row_ptr.empty()) return;
    std::unordered_map<u32, u32> sl; for (u32 s = 0; s < R; ++s) sl[subj[s]] = s;
    for (u32 v = 0; v < N; ++v)
      for (u64 e = row_ptr[v]; e < row_ptr[v + 1]; ++e) { auto it = sl.find(col[e]); if (it != sl.end() && col[e] != v) watch[v] |= (u16)(1u << it->second); }
  }
  int slot_of(u32 nodeid) const { for (u32 s = 0; s < R; ++s) if (subj[s] == nodeid) return (int)s; return -1; }

  // Peer draws: Philox block b yields eight 16-bit draws (low half, then high half of words 0..3);
  // draw i picks neighbour index (h16 * deg) >> 16.  Degrees are limited to 65535 by set_topology.
  static inline u32 draw16(const u32 w[4], u32 i) { const u32 x = w[(i >> 1) & 3]; return (i & 1) ? (x >> 16) : (x & 0xffffu); }
  // Gossip peers (memberlist kRandomNodes: k uniformly random distinct members other than ourselves):
  // m = min(fanout, deg) distinct SLOTS of the node's neighbour list, sampled without replacement by rank —
  // draw k picks rank j = (h16_k * (deg - k)) >> 16 among the slots not chosen yet — then slots that point
  // at the node itself are dropped.  Peers are used in draw order.
  u32 gossip_targets(u32 v, u32 t, u32* out) const {
    u64 r0 = row_ptr[v]; u32 deg = (u32)(row_ptr[v + 1] - r0), nt = 0;
    const u32 m = std::min(cfg.fanout, deg);
    if (!m) return 0;
    u32 w[4];
    philox4x32_10(t, v, 0, DOMAIN_GOSSIP, (u32)cfg.seed, (u32)(cfg.seed >> 32), w);
    u32 chosen[8], nc = 0;                                  // ascending
    for (u32 k = 0; k < m; ++k) {
      u32 j = (draw16(w, k) * (deg - k)) >> 16;
      for (u32 i = 0; i < nc; ++i) if (j >= chosen[i]) ++j;
      u32 pos = nc;                                          // insert keeping the list ascending
      while (pos > 0 && chosen[pos - 1] > j) { chosen[pos] = chosen[pos - 1]; --pos; }
      chosen[pos] = j; ++nc;
      const u32 c = col[r0 + j];
      if (c != v) out[nt++] = c;
    }
    return nt;
  }
  bool probe_target(u32 v, u32 t, u32* out) const {
    u64 r0 = row_ptr[v]; u32 deg = (u32)(row_ptr[v + 1] - r0);
    if (!deg) return false;
    u32 w[4]; philox4x32_10(t, v, 0, DOMAIN_PROBE, (u32)cfg.seed, (u32)(cfg.seed >> 32), w);
    *out = col[r0 + ((draw16(w, 0) * deg) >> 16)];
    return true;
  }

  // One tick.  Nodes are independent within a tick (bulk-synchronous), so the node loop is split into
  // `threads` contiguous id ranges; each range owner gathers its mail, applies it literally and posts the
  // outgoing messages into per-(producer, consumer) boxes.  Results do not depend on `threads`.
  u32 owner_of(u32 v) const { return v / chunk; }
  void step_one() {
    const u32 t = tick;
    const u32 T = (u32)threads;
    if (mail.size() != (size_t)T * T) { mail.assign((size_t)T * T, {}); }
    if (mail_next.size() != (size_t)T * T) { mail_next.assign((size_t)T * T, {}); }
    for (auto& b : mail_next) b.clear();
    if (scratch.size() != T) scratch.assign(T, Scratch{});
    if (exported.size() != T) exported.assign(T, {});
    for (auto& e : exported) e.clear();
    if (ue_n) {
      if (ue_mail.size() != (size_t)T * T) ue_mail.assign((size_t)T * T, {});
      if (ue_mail_next.size() != (size_t)T * T) ue_mail_next.assign((size_t)T * T, {});
      for (auto& b : ue_mail_next) b.clear();
    }
    std::vector<u64> ue_rows((size_t)T * 5, 0);
    std::vector<std::vector<u32>> byz_flagged(T);               // senders judged anomalous by the receivers of each range
    std::vector<u64> byz_rows((size_t)T * 2, 0);
    std::vector<std::pair<u32, u32>> ue_stamps;                 // (event, ltime) stamped this tick; published after the node loop
    std::mutex ue_stamp_mx;
    std::vector<std::vector<Msg>>& next = mail_next;
    // events of this tick
    std::vector<EventB> evs;
    std::unordered_map<u32, EventB> ev_of;
    { auto it = ev_by_tick.find(t); if (it != ev_by_tick.end()) for (auto& e : it->second) { evs.push_back(e); ev_of[e.node] = e; } }
    // ground truth after this tick's operations (what a failed probe observes)
    for (auto& e : evs) {
      int s = slot_of(e.node);
      if (s >= 0) { if (e.op == SERFSIM_OP_FAIL) subj_up[s] = 0; if (e.op == SERFSIM_OP_REJOIN) subj_up[s] = 1; }
    }
    bool any_down = false; for (u32 s = 0; s < R; ++s) any_down |= !subj_up[s];
    const bool reap_now = cfg.reap_interval_ticks && ((t + 1) % cfg.reap_interval_ticks) == 0;
    std::vector<serfsim_tick_row_t> rows(T);

    auto work = [&](u32 c) {
    serfsim_tick_row_t& row = rows[c];
    const u32 v0 = c * chunk, v1 = std::min<u64>(N, (u64)(c + 1) * chunk);
    if (v0 >= v1) return;
    // ---- bucket last tick's messages for my id range by destination (stable counting sort) ----
    Scratch& sx = scratch[c];
    std::vector<u32>& head = sx.head;
    head.assign(v1 - v0 + 1, 0);
    size_t total = 0;
    for (u32 p = 0; p < T; ++p) { for (auto& m : mail[(size_t)p * T + c]) head[m.dst - v0 + 1]++; total += mail[(size_t)p * T + c].size(); }
    for (u32 i = 0; i < v1 - v0; ++i) head[i + 1] += head[i];
    std::vector<Msg>& byd = sx.byd;
    byd.resize(total);
    { std::vector<u32>& pos = sx.pos; pos.assign(head.begin(), head.end() - 1); for (u32 p = 0; p < T; ++p) for (auto& m : mail[(size_t)p * T + c]) byd[pos[m.dst - v0]++] = m; }
    // user-event mail for my id range: destination → arrived tracked events (every copy kept: literal delivery)
    std::unordered_map<u32, std::vector<u32>> ue_in;
    if (ue_n) for (u32 p = 0; p < T; ++p) for (auto& m : ue_mail[(size_t)p * T + c]) ue_in[m.dst].push_back(m.e);
    auto post = [&](const Msg& m) {
      if (m.dst - own_first < own_count) next[(size_t)c * T + owner_of(m.dst)].push_back(m);
      else exported[c].push_back(m);                      // destination lives in another shard
    };
    for (u32 v = std::max(v0, own_first); v < std::min<u64>(v1, (u64)own_first + own_count); ++v) {
      NodeB& nd = node[v];
      const bool up_r = nd.up;
      const EventB* ev = nullptr;
      if (!ev_of.empty()) { auto it = ev_of.find(v); if (it != ev_of.end()) ev = &it->second; }   // at most one op per (node, tick)
      bool up_s = up_r;
      if (ev && ev->op == SERFSIM_OP_FAIL) up_s = false;
      if (ev && ev->op == SERFSIM_OP_REJOIN) up_s = true;
      u32 targets[8], nt = 0; bool have_targets = false;
      u32 ptarget = 0; bool have_probe = false;
      const u32 wmask = watch.empty() ? 0u : watch[v];
      if (up_s && wmask && cfg.probe_interval_ticks && any_down && ((t + v) % cfg.probe_interval_ticks) == 0)
        have_probe = probe_target(v, t, &ptarget);
      u32 max_tx = 0;
      for (u32 s = 0; s < R; ++s) {
        View& r = at(s, v);
        const bool self = (subj[s] == v);
        // ---------------- Phase R: receive / state merge ----------------
        if (up_r) {
          const View before = r;
          // this node's mail for slot s, in canonical order: memberlist messages, then leave intents ascending
          // (ltime, src), then join intents ascending — sorted once per node (first slot), then walked per slot
          if (s == 0 && head[v - v0 + 1] - head[v - v0] > 1)
            std::sort(byd.begin() + head[v - v0], byd.begin() + head[v - v0 + 1], [](const Msg& a, const Msg& b) {
              if (a.slot != b.slot) return a.slot < b.slot;
              const int ka = a.kind == 2 ? 0 : a.kind == 0 ? 1 : 2, kb = b.kind == 2 ? 0 : b.kind == 0 ? 1 : 2;
              if (ka != kb) return ka < kb;
              if (a.val != b.val) return a.val < b.val;
              if (a.kind == 0 && a.prune != b.prune) return a.prune > b.prune;   // leave intents of equal Lamport time: the pruning one first (DESIGN rule P-1)
              return a.src < b.src;
            });
          u32 i0 = head[v - v0];
          const u32 i1 = head[v - v0 + 1];
          while (i0 < i1 && byd[i0].slot < s) ++i0;
          // memberlist: only the greatest (incarnation, kind, from) message is delivered per tick (rule ML-1)
          {
            u32 key = 0; bool any = false;
            for (; i0 < i1 && byd[i0].slot == s && byd[i0].kind == 2; ++i0) { key = std::max(key, byd[i0].val); any = true; }
            if (any) {
              u32 inc = key >> 6, kind = (key >> 4) & 3, fromh = key & 15;
              if (kind == ML_ALIVE) v_ml_alive(r, inc, self, cx);
              else if (kind == ML_SUSPECT) v_ml_suspect(r, inc, fromh, t, self, cx);
              else v_ml_dead(r, inc, kind == ML_LEFT, t, self, cx);
            }
          }
          // serf intents, one message at a time
          bool refute = false;
          {
            // rule P-1: of the leave intents of one tick only the greatest (ltime, then non-pruning over pruning) keeps its prune
            // flag — every copy of it does (gossip delivers the same intent from several peers; the first copy is accepted)
            u32 j = i0;
            while (j < i1 && byd[j].slot == s && byd[j].kind == 0) ++j;
            for (u32 i = i0; i < j; ++i) {
              // rule P-2: further copies of one intent within a tick are dropped.  Without prune a second copy is rejected anyway
              // (ltime <= status_time, or no newer than the buffered intent); after a pruning one the member is unknown again
              // and the reference would buffer the second copy as a fresh intent — the reduced inbox cannot count copies.
              if (i > i0 && byd[i].val == byd[i - 1].val && byd[i].prune == byd[i - 1].prune) continue;
              const bool greatest = byd[i].val == byd[j - 1].val && byd[i].prune == byd[j - 1].prune;
              witness32(nd.clock, byd[i].val);
              v_leave_intent(r, byd[i].val, self, nd.sstate, &refute, cx, true, greatest && byd[i].prune);
            }
            i0 = j;
          }
          for (; i0 < i1 && byd[i0].slot == s && byd[i0].kind == 1; ++i0) { witness32(nd.clock, byd[i0].val); v_join_intent(r, byd[i0].val, cx); }
          if (refute) {                                   // base.rs:1470-1480 → broadcast_join(clock.time()), base.rs:381-397
            u32 T = nd.clock; witness32(nd.clock, T);
            v_join_intent(r, T, cx);
            r.qjoin = T; r.txj = (u8)cx.limit;
          }
          if (!known(r) && r.status != TY_NONE && (r.status != before.status || r.st != before.st)) r.leave_tick = t + 1;   // NodeIntent.wall_time (types/member.rs:32)
          if (memcmp(&before, &r, sizeof(View)) != 0) row.changed++;
        }
        // ---------------- Phase E: host operation ----------------
        if (ev) {
          const u32 op = ev->op;
          if (op == SERFSIM_OP_FAIL && self) { /* state kept; node simply stops */ }
          if (op == SERFSIM_OP_REJOIN && self && !up_r) {
            r.inc += 1; set_ml(r, ML_ALIVE, 0); r.txm = (u8)cx.limit; r.deadline = 0; r.mask = 0;
            nd.sstate = SS_ALIVE;
            v_node_join(r);
          }
          if (((op == SERFSIM_OP_JOIN && up_r) || (op == SERFSIM_OP_REJOIN && !up_r)) && self) {   // api.rs:339-342 → base.rs:381-397
            u32 T = nd.clock; witness32(nd.clock, T);
            v_join_intent(r, T, cx);
            r.qjoin = T; r.txj = (u8)cx.limit;
          }
          if (op == SERFSIM_OP_LEAVE && up_r && self && nd.sstate == SS_ALIVE) {                  // api.rs:422-449
            nd.sstate = SS_LEAVING;
            u32 T = nd.clock; nd.clock += 1;
            bool refute = false;
            v_leave_intent(r, T, true, nd.sstate, &refute, cx);
            r.qleave = T; r.txl = (u8)cx.limit; r.flags &= (u8)~2u;
          }
          if ((op == SERFSIM_OP_FORCE_LEAVE || op == SERFSIM_OP_FORCE_LEAVE_PRUNE) && up_r && ev->slot == s) {   // base.rs:454-480; api.rs:500-515
            const bool prune = op == SERFSIM_OP_FORCE_LEAVE_PRUNE;
            u32 T = nd.clock; witness32(nd.clock, T);
            bool refute = false;
            v_leave_intent(r, T, self, nd.sstate, &refute, cx, true, prune);
            r.qleave = T; r.txl = (u8)cx.limit; r.flags = (u8)((r.flags & ~2u) | (prune ? 2u : 0u));   // queued whatever the handler said (base.rs:466-476)
            if (refute) { u32 T2 = nd.clock; witness32(nd.clock, T2); v_join_intent(r, T2, cx); r.qjoin = T2; r.txj = (u8)cx.limit; }
          }
        }
        if (ev && !known(r) && r.status != TY_NONE && r.leave_tick == 0) r.leave_tick = t + 1;
        if (up_s) {
          // ---------------- Phase T: reaper (serf/base.rs:483-610), suspicion timer, probe ----------------
          if (reap_now) {
            const u32 age = r.leave_tick ? (t + 1 - r.leave_tick) : 0;
            if (known(r) && r.leave_tick &&
                ((r.status == ST_LEFT && age > cfg.tombstone_timeout_ticks) || (r.status == ST_FAILED && age > cfg.reconnect_timeout_ticks))) {
              r.flags &= ~1; r.status = TY_NONE; r.st = 0; r.leave_tick = 0;          // erase_node! :499-519 (the view forgets the member)
            } else if (!known(r) && r.status != TY_NONE && r.leave_tick && age > cfg.recent_intent_timeout_ticks) {
              r.status = TY_NONE; r.st = 0; r.leave_tick = 0;                          // reap_intents :1817-1822
            }
          }
          if (ml_state(r) == ML_SUSPECT && r.deadline != 0 && t >= r.deadline) v_ml_dead(r, r.inc, false, t, false, cx);
          if (have_probe && !self && ptarget == subj[s] && !subj_up[s]) {
            u8 st = ml_state(r);
            if (st == ML_ALIVE || st == ML_SUSPECT) {
              bool starts = (st == ML_ALIVE);
              v_ml_suspect(r, r.inc, from_hash(v), t, false, cx);
              if (starts) row.suspects++;
            }
          }
          // ---------------- Phase S: gossip send ----------------
          if (r.txl | r.txj | r.txm) {
            if (!have_targets) { nt = gossip_targets(v, t, targets); have_targets = true; }
            for (u32 k = 0; k < nt; ++k) {
              u32 cnt = 0;
              if (r.txl > k) { post(Msg{targets[k], v, r.qleave, (u8)s, 0, 0, (u8)((r.flags >> 1) & 1u)}); ++cnt; }
              if (r.txj > k) { post(Msg{targets[k], v, r.qjoin, (u8)s, 1}); ++cnt; }
              if (r.txm > k) { post(Msg{targets[k], v, ml_key(r), (u8)s, 2}); ++cnt; }
              if (cnt) { row.edge_updates++; row.messages += cnt; }
            }
            max_tx = std::max(max_tx, (u32)std::max(r.txl, std::max(r.txj, r.txm)));
            r.txl -= (u8)std::min<u32>(r.txl, nt); r.txj -= (u8)std::min<u32>(r.txj, nt); r.txm -= (u8)std::min<u32>(r.txm, nt);
          }
          // Serf::leave: once our own leave intent is out, memberlist.leave() → dead{node == from}  (api.rs:451-476)
          if (self && nd.sstate == SS_LEAVING && r.txl == 0 && ml_state(r) == ML_ALIVE) {
            set_ml(r, ML_LEFT, 0); r.txm = (u8)cx.limit; nd.sstate = SS_LEFT;
          }
          // pending?
          bool pend = (r.txl | r.txj | r.txm) || ml_state(r) == ML_SUSPECT ||
                      (cfg.probe_interval_ticks && !subj_up[s] && !self && ((wmask >> s) & 1) && ml_state(r) == ML_ALIVE);   // a watcher that has not noticed yet
          if (pend) row.pending++;
          // byzantine injector: a stale copy of the END-of-tick view goes to this tick's gossip peers, budgets or not
          if (byz_n && byz[v] && known(r)) {
            if (!have_targets) { nt = gossip_targets(v, t, targets); have_targets = true; }
            u8 kind; u32 lt, key;
            byz_stale(r, byz_delta, &kind, &lt, &key);
            for (u32 k = 0; k < nt; ++k) {
              post(Msg{targets[k], v, lt, (u8)s, kind, 1});
              post(Msg{targets[k], v, key, (u8)s, 2, 1});
              byz_rows[(size_t)c * 2] += 2; byz_rows[(size_t)c * 2 + 1] += 1;
            }
          }
        }
      }
      // ---------------- user events: receive, originate, send (serf/base.rs:750-837, serf/api.rs:241-299) ----------------
      if (ue_n) {
        UeNodeB& un = uen[v];
        u64* ur = &ue_rows[(size_t)c * 5];
        auto itin = ue_in.empty() ? ue_in.end() : ue_in.find(v);
        if (up_r) {
          if (itin != ue_in.end()) {
            std::vector<u32>& arr = itin->second;
            std::sort(arr.begin(), arr.end());                                  // canonical order: ascending tracked-event index
            for (size_t i = 0; i < arr.size(); ++i) {
              const u32 e = arr[i];
              const int oc = ue_handle(un, e, ue_ltime[e]);
              if (oc == 0) un.tx[e] = (u8)cx.limit;                             // true → re-queued with a fresh budget (delegate.rs:293-300)
              if (i == 0 || arr[i - 1] != e) {                                  // counters per (node, tick, event): copies 2..k are always duplicates
                if (oc == 0) { ur[2]++; row.changed++; } else if (oc == 1) ur[3]++; else ur[4]++;
              }
            }
          }
          if (ev && ev->op == SERFSIM_OP_USER_EVENT) {                          // Serf::user_event, serf/api.rs:241-299
            const u32 e = ev->slot;
            const u32 L = un.clock;                                             // :264 ltime = event_clock.time()
            un.clock += 1;                                                      // :285 increment
            const int oc = ue_handle(un, e, L);                                 // :288 handled locally, result ignored
            un.tx[e] = (u8)cx.limit;                                            // :290-297 queued unconditionally
            if (oc == 0) { ur[2]++; row.changed++; } else if (oc == 1) ur[3]++; else ur[4]++;
            std::lock_guard<std::mutex> g(ue_stamp_mx);
            ue_stamps.push_back({e, L});
          }
        }
        if (up_s) {
          bool any = false; for (u32 e = 0; e < ue_n; ++e) any |= un.tx[e] != 0;
          if (any) {
            u32 utg[8]; const u32 unt = gossip_targets(v, t, utg);              // same packet, same peers as the intents
            for (u32 k = 0; k < unt; ++k) {
              u32 cnt = 0;
              for (u32 e = 0; e < ue_n; ++e) if (un.tx[e] > k) { ue_mail_next[(size_t)c * T + owner_of(utg[k])].push_back(UeMsg{utg[k], e}); ++cnt; }
              if (cnt) { row.edge_updates++; row.messages += cnt; ur[0] += cnt; ur[1]++; }
            }
            for (u32 e = 0; e < ue_n; ++e) un.tx[e] -= (u8)std::min<u32>(un.tx[e], unt);
          }
        }
        if (up_s) for (u32 e = 0; e < ue_n; ++e) if (un.tx[e]) row.pending++;      // a crashed node's queue is frozen, not pending
      }
      if (ev) { row.events++; }
      nd.up = up_s;
      row.packets += std::min(nt, max_tx);
    }
    };
    if (T == 1) work(0);
    else run_workers(T, work);
    mail.swap(mail_next);
    if (byz_n) {
      // Verdicts: every stale entry posted this tick is judged against its receiver's view as it stands when the node loop
      // of this tick is over — the state the packet meets on arrival, except for what a push-pull round of this very tick
      // still merges afterwards (the round runs after the verdicts; the device does the same).  A receiver that is down
      // at that point never sees the packet.  `mail` holds this tick's messages after the swap above.
      for (auto& box : mail)
        for (const Msg& m : box) {
          if (!m.byz || !node[m.dst].up) continue;
          if (byz_judge(at(m.slot, m.dst), m.kind, m.val, byz_delta)) byz_flagged[0].push_back(m.src);
        }
      for (u32 c = 0; c < T; ++c) {
        for (u32 src : byz_flagged[c]) if (!anomaly[src]) { anomaly[src] = 1; byz_tot[2]++; }
        byz_tot[0] += byz_rows[(size_t)c * 2]; byz_tot[1] += byz_rows[(size_t)c * 2 + 1];
      }
    }
    if (ue_n) {
      ue_mail.swap(ue_mail_next);
      for (auto& st : ue_stamps) { ue_ltime[st.first] = st.second; ue_stamped |= 1u << st.first; }   // visible to receivers from the next tick on
      // cluster runs need one Lamport time per ring slot (the packed record derives the slot's ltime from its events)
      for (u32 a = 0; a < ue_n; ++a) for (u32 b = a + 1; b < ue_n; ++b)
        if (((ue_stamped >> a) & 1) && ((ue_stamped >> b) & 1) && ue_ltime[a] % 512 == ue_ltime[b] % 512 && ue_ltime[a] != ue_ltime[b]) ue_alias_error = true;
      for (u32 c = 0; c < T; ++c) for (int i = 0; i < 5; ++i) ue_tot[i] += ue_rows[(size_t)c * 5 + i];
    }
    // ---------------- anti-entropy round: memberlist push-pull + SerfDelegate::merge_remote_state ----------------
    // (serf/delegate.rs:386-554; memberlist mergeState [external]).  Every push_pull_interval ticks each up node
    // pulls the end-of-tick state of ONE random neighbour and merges it: clock witness(ltime-1); per subject the
    // memberlist state (alive → aliveNode, suspect/dead → suspectNode{from = self}, left → deadNode{from = node}),
    // then serf's view: a Left member → leave intent at status_ltime + 1, any other known member → join intent
    // at status_ltime — results discarded, i.e. nothing is re-queued (delegate.rs:495-523).
    const u32 pp = (u32)std::max(0, cfg.push_pull_interval_ticks);
    if (pp && (t + 1) % pp == 0 && own_count == N) {
      const std::vector<View> srec = rec;
      const std::vector<NodeB> snode = node;
      const std::vector<UeNodeB> suen = uen;                    // PushPullMessage.event_ltime / .events of every node (delegate.rs:386-425)
      std::vector<u64> pp_ue((size_t)T * 3, 0);
      auto ppwork = [&](u32 c) {
        serfsim_tick_row_t& row = rows[c];
        const u32 v0 = c * chunk, v1 = std::min<u64>(N, (u64)(c + 1) * chunk);
        for (u32 v = v0; v < v1; ++v) {
          NodeB& nd = node[v];
          if (!snode[v].up) continue;
          const u64 r0 = row_ptr[v]; const u32 deg = (u32)(row_ptr[v + 1] - r0);
          if (!deg) continue;
          u32 w[4]; philox4x32_10(t, v, 0, DOMAIN_PUSHPULL, (u32)cfg.seed, (u32)(cfg.seed >> 32), w);
          const u32 u = col[r0 + ((draw16(w, 0) * deg) >> 16)];
          if (u == v || !snode[u].up) continue;
          if (snode[u].clock > 0) witness32(nd.clock, snode[u].clock - 1);            // delegate.rs:466-468
          if (ue_n) {                                                                  // delegate.rs:469-474 and 539-552
            UeNodeB& un = uen[v];
            const UeNodeB& pu = suen[u];
            if (pu.clock > 0) witness32(un.clock, pu.clock - 1);
            for (auto& kv : pu.ring)                                                   // the partner's buffer in ring-index order
              for (u32 e : kv.second.events) {
                const int oc = ue_handle(un, e, kv.second.ltime);                      // handled, result discarded: nothing is re-queued
                if (oc == 0) { pp_ue[(size_t)c * 3]++; row.changed++; } else if (oc == 1) pp_ue[(size_t)c * 3 + 1]++; else pp_ue[(size_t)c * 3 + 2]++;
              }
          }
          for (u32 s = 0; s < R; ++s) {
            View& r = at(s, v);
            const View before = r;
            const View& q = srec[(size_t)s * N + u];
            const bool self = (subj[s] == v);
            const u32 wmask = watch.empty() ? 0u : watch[v];
            const bool was_pending = (r.txl | r.txj | r.txm) || ml_state(r) == ML_SUSPECT ||
                                     (cfg.probe_interval_ticks && !subj_up[s] && !self && ((wmask >> s) & 1) && ml_state(r) == ML_ALIVE);
            if (known(q)) {
              const u8 qs = ml_state(q);
              if (qs == ML_ALIVE) v_ml_alive(r, q.inc, self, cx);
              else if (qs == ML_LEFT) v_ml_dead(r, q.inc, true, t, self, cx);
              else v_ml_suspect(r, q.inc, from_hash(v), t, self, cx);
              bool refute = false;
              if (q.status == ST_LEFT) { witness32(nd.clock, q.st + 1); v_leave_intent(r, q.st + 1, self, nd.sstate, &refute, cx, false); }
              else { witness32(nd.clock, q.st); v_join_intent(r, q.st, cx, false); }
              if (refute) { u32 T2 = nd.clock; witness32(nd.clock, T2); v_join_intent(r, T2, cx); r.qjoin = T2; r.txj = (u8)cx.limit; }
              if (!known(r) && r.status != TY_NONE && (r.status != before.status || r.st != before.st)) r.leave_tick = t + 1;
            }
            { View a = before, b = r; a.st = b.st = 0;       // status_time creeps by design (leave at status_ltime + 1): not a change
              if (memcmp(&a, &b, sizeof(View)) != 0) row.changed++; }
            const bool now_pending = (r.txl | r.txj | r.txm) || ml_state(r) == ML_SUSPECT ||
                                     (cfg.probe_interval_ticks && !subj_up[s] && !self && ((wmask >> s) & 1) && ml_state(r) == ML_ALIVE);
            if (now_pending && !was_pending) row.pending++;
            if (!now_pending && was_pending) row.pending--;
          }
        }
      };
      if (T == 1) ppwork(0);
      else run_workers(T, ppwork);
      for (u32 c = 0; c < T; ++c) { ue_tot[2] += pp_ue[(size_t)c * 3]; ue_tot[3] += pp_ue[(size_t)c * 3 + 1]; ue_tot[4] += pp_ue[(size_t)c * 3 + 2]; }
    }
    serfsim_tick_row_t row{};
    for (auto& r : rows) { row.packets += r.packets; row.edge_updates += r.edge_updates; row.messages += r.messages; row.changed += r.changed;
                           row.pending += r.pending; row.events += r.events; row.suspects += r.suspects; }
    if (cfg.trace) row.hash = state_hash();
    trace.push_back(row);
    tot_events += row.events;
    ++tick;
  }

  u64 state_hash() const {                    // additive over nodes: shard hashes sum to the global hash
    u64 h = 0;
    const u32 a = own_first, b = own_first + own_count;
    for (u32 s = 0; s < R; ++s)
      for (u32 v = a; v < b; ++v) {
        u64 w[4]; memcpy(w, &rec[(size_t)s * N + v], 32);
        u64 idx = (u64)s * N + v;
        h += mix64(w[0] ^ mix64(w[1] ^ mix64(w[2] ^ mix64(w[3] ^ mix64(idx + 0x9e3779b97f4a7c15ULL)))));
      }
    for (u32 v = a; v < b; ++v) {
      u64 w = (u64)node[v].clock | ((u64)node[v].up << 32) | ((u64)node[v].sstate << 40);
      h += mix64(w ^ mix64((u64)R * N + v + 0x9e3779b97f4a7c15ULL));
    }
    if (ue_n) for (u32 v = a; v < b; ++v) {
      u32 w[4]; ue_record(v, w);
      h += mix64((((u64)w[1] << 32) | w[0]) ^ mix64((((u64)w[3] << 32) | w[2]) ^ mix64((u64)(R + 1) * N + v + 0x9e3779b97f4a7c15ULL)));
    }
    return h;
  }
  bool future_events() const { return any_event && max_event_tick >= tick; }
};

// =====================================================================================
// C ABI (ctypes)
// =====================================================================================
static thread_local std::string g_err;
#define ORC extern "C" __attribute__((visibility("default")))

// ---- Part A ----
ORC void* ref_node_new(u64 self_id, u32 retransmit_mult) { auto* n = new RefNode(self_id); n->retransmit_mult = retransmit_mult; return n; }
ORC void ref_node_free(void* p) { delete (RefNode*)p; }
ORC u64 ref_clock_time(void* p, int which) { auto* n = (RefNode*)p; return (which == 0 ? n->clock : which == 1 ? n->event_clock : n->query_clock).time(); }
ORC u64 ref_clock_increment(void* p) { return ((RefNode*)p)->clock.increment(); }
ORC void ref_clock_witness(void* p, u64 t) { ((RefNode*)p)->clock.witness(t); }
ORC u64 lamport_new_time(void) { LamportClock c; return c.time(); }
ORC void ref_set_serf_state(void* p, int s) { ((RefNode*)p)->serf_state = (u8)s; }
ORC int ref_get_serf_state(void* p) { return ((RefNode*)p)->serf_state; }
ORC void ref_insert_member(void* p, u64 id, int status, u64 status_time) { ((RefNode*)p)->states[id] = MemberStateA{(u8)status, status_time, false, 0}; }
ORC int ref_member_get(void* p, u64 id, u8* status, u64* status_time) {
  auto* n = (RefNode*)p; auto it = n->states.find(id);
  if (it == n->states.end()) return 0;
  *status = it->second.status; *status_time = it->second.status_time; return 1;
}
ORC u64 ref_num_members(void* p) { return ((RefNode*)p)->states.size(); }
ORC int ref_handle_node_join_intent(void* p, u64 ltime, u64 id) { return ((RefNode*)p)->handle_node_join_intent(ltime, id); }
ORC int ref_handle_node_leave_intent(void* p, u64 ltime, u64 id, int prune) { return ((RefNode*)p)->handle_node_leave_intent(ltime, id, prune != 0); }
ORC void ref_run_detached(void* p) { ((RefNode*)p)->run_detached(); }
ORC u32 ref_refutes(void* p) { return ((RefNode*)p)->refutes; }
ORC void ref_handle_node_join(void* p, u64 id) { ((RefNode*)p)->handle_node_join(id); }
ORC void ref_handle_node_leave(void* p, u64 id, int64_t now_ms) { ((RefNode*)p)->handle_node_leave(id, now_ms); }
ORC int ref_upsert_intent(void* p, u64 id, int ty, u64 ltime, int64_t wall_ms) { return ((RefNode*)p)->upsert_intent(id, (u8)ty, ltime, wall_ms); }
ORC int ref_recent_intent(void* p, u64 id, int ty, u64* ltime) { return ((RefNode*)p)->recent_intent(id, (u8)ty, ltime); }
ORC void ref_reap_intents(void* p, int64_t now_ms, int64_t timeout_ms) { ((RefNode*)p)->reap_intents(now_ms, timeout_ms); }
ORC void ref_merge_remote_state(void* p, u64 pp_ltime, const u64* ids, const u64* ltimes, u32 n, const u64* left, u32 n_left, u64 event_ltime, u64 query_ltime) {
  ((RefNode*)p)->merge_remote_state(pp_ltime, ids, ltimes, n, left, n_left, event_ltime, query_ltime);
}
ORC u32 ref_local_state(void* p, u64* pp_ltime, u64* ids, u64* ltimes, u32 cap, u64* left, u32 cap_left, u32* n_left, u64* event_ltime, u64* query_ltime) {
  return ((RefNode*)p)->local_state(pp_ltime, ids, ltimes, cap, left, cap_left, n_left, event_ltime, query_ltime);
}
ORC int ref_handle_user_event(void* p, u64 ltime, const char* name, const char* payload) { return ((RefNode*)p)->handle_user_event(ltime, name, payload); }
ORC void ref_event_clock_witness(void* p, u64 t) { ((RefNode*)p)->event_clock.witness(t); }
ORC u32 ref_user_event_count(void* p) { return (u32)((RefNode*)p)->user_events_out.size(); }
ORC const char* ref_user_event_get(void* p, u32 i, int payload) { auto& e = ((RefNode*)p)->user_events_out[i]; return payload ? e.second.c_str() : e.first.c_str(); }
ORC int ref_event_buffer_has(void* p, u64 ltime) { auto* n = (RefNode*)p; auto& sl = n->event_buffer[(size_t)(ltime % n->event_buffer.size())]; return sl.first && sl.second.ltime == ltime; }
ORC void ref_api_join(void* p) { ((RefNode*)p)->api_join(); }
ORC int ref_api_leave(void* p) { return ((RefNode*)p)->api_leave(); }
ORC void ref_api_force_leave(void* p, u64 id, int prune) { ((RefNode*)p)->api_force_leave(id, prune != 0); }
ORC u32 ref_queue_len(void* p) { return (u32)((RefNode*)p)->broadcasts.size(); }
ORC int ref_queue_get(void* p, u32 i, u8* ty, u64* ltime, u64* id, u32* transmits) {
  auto* n = (RefNode*)p; if (i >= n->broadcasts.size()) return 0;
  auto& q = n->broadcasts[i]; *ty = q.ty; *ltime = q.ltime; *id = q.id; *transmits = q.transmits; return 1;
}
ORC u32 ref_get_broadcasts(void* p, u32 byte_limit, u32 overhead, u8* ty, u64* lt, u64* id, u32 cap) { return ((RefNode*)p)->get_broadcasts(byte_limit, overhead, ty, lt, id, cap); }
ORC void ref_remove_old_member(void* p, int failed_list, u64 id) { auto* n = (RefNode*)p; RefNode::remove_old_member(failed_list ? n->failed_members : n->left_members, id); }
ORC u32 ref_left_count(void* p) { return (u32)((RefNode*)p)->left_members.size(); }
ORC u32 ref_failed_count(void* p) { return (u32)((RefNode*)p)->failed_members.size(); }
ORC void ref_push_left(void* p, u64 id, int status, u64 status_time, int64_t leave_time_ms) { ((RefNode*)p)->left_members.push_back({id, MemberStateA{(u8)status, status_time, true, leave_time_ms}}); }
ORC void ref_push_failed(void* p, u64 id, int status, u64 status_time, int64_t leave_time_ms) { ((RefNode*)p)->failed_members.push_back({id, MemberStateA{(u8)status, status_time, true, leave_time_ms}}); }
ORC void ref_reap(void* p, int64_t now_ms, int64_t reconnect_ms, int64_t tombstone_ms, int64_t intent_ms) { ((RefNode*)p)->reap(now_ms, reconnect_ms, tombstone_ms, intent_ms); }
ORC u32 ref_event_count(void* p) { return (u32)((RefNode*)p)->events.size(); }
ORC int ref_event_get(void* p, u32 i, u32* ty, u64* id) { auto* n = (RefNode*)p; if (i >= n->events.size()) return 0; *ty = n->events[i].first; *id = n->events[i].second; return 1; }
ORC u64 ref_get_queue_max(u64 max_depth, u64 min_depth, u64 members) { return get_queue_max(max_depth, min_depth, members); }

// ---- shared primitives ----
ORC void oracle_philox4x32_10(const u32 ctr[4], const u32 key[2], u32 out[4]) { philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out); }
ORC u32 oracle_retransmit_limit(u32 mult, u64 n) { return retransmit_limit(mult, n); }
ORC u32 oracle_suspicion_table(u32 susp_mult, u32 max_mult, u32 probe_ticks, u32 tick_ms, u64 n, u32* out, u32 cap) {
  auto t = suspicion_table(susp_mult, max_mult, probe_ticks, tick_ms, n);
  for (u32 i = 0; i < t.size() && i < cap; ++i) out[i] = t[i];
  return (u32)t.size();
}
ORC u64 oracle_mix64(u64 x) { return mix64(x); }
ORC u32 oracle_from_hash(u32 node) { return from_hash(node); }

// view-rule hooks: lets tests drive the packed-record rules against RefNode (Part A ≡ Part B)
ORC void oracle_view_init(void* rec32, int known_, int status, u32 st) { View r{}; r.flags = known_ ? 1 : 0; r.status = (u8)status; r.st = st; r.inc = known_ ? 1 : 0; set_ml(r, known_ ? ML_ALIVE : ML_DEAD, 0); memcpy(rec32, &r, 32); }
ORC int oracle_view_join_intent(void* rec32, u32 lt, u32 limit) { RuleCtx cx{limit, 0, nullptr}; View r; memcpy(&r, rec32, 32); bool a = v_join_intent(r, lt, cx); memcpy(rec32, &r, 32); return a; }
ORC int oracle_view_leave_intent(void* rec32, u32 lt, int self, int sstate, int* refute, u32 limit) { RuleCtx cx{limit, 0, nullptr}; View r; memcpy(&r, rec32, 32); bool rf = false; bool a = v_leave_intent(r, lt, self != 0, (u8)sstate, &rf, cx); memcpy(rec32, &r, 32); *refute = rf; return a; }
ORC void oracle_view_node_join(void* rec32) { View r; memcpy(&r, rec32, 32); v_node_join(r); memcpy(rec32, &r, 32); }
ORC void oracle_view_node_leave(void* rec32, u32 tick) { View r; memcpy(&r, rec32, 32); v_node_leave(r, tick); memcpy(rec32, &r, 32); }

// ---- Part B: same shapes as the serfsim_* ABI so one test driver serves both ----
ORC int oracle_sim_create(const serfsim_config_t* cfg, void** out) {
  if (!cfg || !out || cfg->n_nodes == 0 || cfg->slots == 0 || cfg->slots > 16 || cfg->fanout == 0 || cfg->fanout > 8) { g_err = "bad config"; return SERFSIM_E_INVAL; }
  auto* s = new TickSim(); s->cfg = *cfg; s->N = cfg->n_nodes; s->R = cfg->slots;
  s->subj.resize(s->R); for (u32 i = 0; i < s->R; ++i) s->subj[i] = i;
  s->init_tables();
  if (s->cx.limit > 255) { delete s; g_err = "retransmit limit > 255"; return SERFSIM_E_INVAL; }
  s->reset(cfg->seed);
  *out = s; return 0;
}
ORC void oracle_sim_destroy(void* p) { delete (TickSim*)p; }
ORC const char* oracle_last_error(void) { return g_err.c_str(); }
ORC int oracle_sim_set_topology_csr(void* p, const u64* row_ptr, const u32* col_idx) {
  auto* s = (TickSim*)p; s->row_ptr.assign(row_ptr, row_ptr + s->N + 1); s->col.assign(col_idx, col_idx + row_ptr[s->N]);
  for (u32 c : s->col) if (c >= s->N) { g_err = "col_idx out of range"; return SERFSIM_E_INVAL; }
  for (u32 v = 0; v < s->N; ++v) if (row_ptr[v + 1] - row_ptr[v] > 65535) { g_err = "node degree > 65535"; return SERFSIM_E_INVAL; }
  s->compute_watch();
  return 0;
}
ORC int oracle_sim_set_subjects(void* p, const u32* subjects) {
  auto* s = (TickSim*)p;
  for (u32 i = 0; i < s->R; ++i) { if (subjects[i] >= s->N) return SERFSIM_E_INVAL; for (u32 j = 0; j < i; ++j) if (subjects[j] == subjects[i]) return SERFSIM_E_INVAL; }
  s->subj.assign(subjects, subjects + s->R); s->compute_watch(); return 0;
}
ORC int oracle_sim_reset(void* p, u64 seed) { ((TickSim*)p)->reset(seed); return 0; }
ORC int oracle_sim_inject(void* p, u32 tick, u32 op, u32 node, u32 slot) {
  auto* s = (TickSim*)p;
  if (tick < s->tick || node >= s->N || op < 1 || op > SERFSIM_OP_FORCE_LEAVE_PRUNE) { g_err = "bad inject"; return SERFSIM_E_INVAL; }
  if (op == SERFSIM_OP_USER_EVENT) {
    if (slot >= s->ue_n) { g_err = "user event index out of range"; return SERFSIM_E_INVAL; }
    if ((s->ue_injected >> slot) & 1u) { g_err = "a tracked user event can be injected once"; return SERFSIM_E_INVAL; }
  }
  if (op == SERFSIM_OP_FORCE_LEAVE || op == SERFSIM_OP_FORCE_LEAVE_PRUNE) { if (slot >= s->R) return SERFSIM_E_INVAL; }
  else if ((op == SERFSIM_OP_JOIN || op == SERFSIM_OP_LEAVE) && s->slot_of(node) < 0) { g_err = "join/leave origin must be a tracked subject"; return SERFSIM_E_INVAL; }
  if (!s->event_keys.insert(((u64)tick << 32) | node).second) { g_err = "one operation per node per tick"; return SERFSIM_E_INVAL; }
  if (op == SERFSIM_OP_USER_EVENT) s->ue_injected |= 1u << slot;
  s->events.push_back(EventB{tick, op, node, slot}); s->ev_by_tick[tick].push_back(EventB{tick, op, node, slot});
  s->max_event_tick = s->any_event ? std::max(s->max_event_tick, tick) : tick; s->any_event = true;
  return 0;
}
static int ue_alias_check(TickSim* s) { if (s->ue_alias_error) { g_err = "user events: two tracked events share a ring slot with different Lamport times (not supported in cluster runs)"; return SERFSIM_E_INVAL; } return 0; }
ORC int oracle_sim_step(void* p, u32 n) { auto* s = (TickSim*)p; if (s->row_ptr.empty()) { g_err = "no topology"; return SERFSIM_E_INVAL; } for (u32 i = 0; i < n; ++i) s->step_one(); return ue_alias_check(s); }
ORC int oracle_sim_run_until_converged(void* p, u32 max_ticks, u32* ticks_out) {
  auto* s = (TickSim*)p; if (s->row_ptr.empty()) return SERFSIM_E_INVAL;
  for (u32 i = 0; i < max_ticks; ++i) {
    s->step_one();
    auto& r = s->trace.back();
    const u32 pp = (u32)std::max(0, s->cfg.push_pull_interval_ticks);
    // with anti-entropy on, convergence additionally needs a push-pull round that changed nothing but Lamport
    // times (status_time keeps creeping: the reference re-sends a Left member as "leave at status_ltime + 1",
    // serf/delegate.rs:495-510, so the round's `changed` counter ignores status_time)
    const bool pp_ok = !pp || ((s->tick % pp) == 0 && r.changed == 0);
    // with byzantine injectors stale entries are in flight forever: quiescent = no honest traffic AND nothing merged this tick
    const bool byz_ok = !s->byz_n || r.changed == 0;
    if (s->ue_alias_error) return ue_alias_check(s);
    if (r.pending == 0 && r.edge_updates == 0 && !s->future_events() && pp_ok && byz_ok) { if (ticks_out) *ticks_out = s->tick - 1; return 0; }
  }
  if (ticks_out) *ticks_out = s->tick;
  return 1;
}
ORC int oracle_sim_member_status(void* p, u32 slot, u8* out) { auto* s = (TickSim*)p; if (slot >= s->R) return SERFSIM_E_INVAL; for (u32 v = 0; v < s->N; ++v) { View& r = s->at(slot, v); out[v] = known(r) ? r.status : (u8)ST_NONE; } return 0; }
ORC int oracle_sim_status_ltime(void* p, u32 slot, u64* out) { auto* s = (TickSim*)p; if (slot >= s->R) return SERFSIM_E_INVAL; for (u32 v = 0; v < s->N; ++v) { View& r = s->at(slot, v); out[v] = known(r) ? r.st : 0; } return 0; }
ORC int oracle_sim_status_ltime_u32(void* p, u32 slot, u32* out) { auto* s = (TickSim*)p; if (slot >= s->R) return SERFSIM_E_INVAL; for (u32 v = 0; v < s->N; ++v) { View& r = s->at(slot, v); out[v] = known(r) ? r.st : 0; } return 0; }
ORC int oracle_sim_lamport_time_u32(void* p, u32* out) { auto* s = (TickSim*)p; for (u32 v = 0; v < s->N; ++v) out[v] = s->node[v].clock; return 0; }
ORC int oracle_sim_lamport_time(void* p, u64* out) { auto* s = (TickSim*)p; for (u32 v = 0; v < s->N; ++v) out[v] = s->node[v].clock; return 0; }
ORC int oracle_sim_incarnation(void* p, u32 slot, u32* out) { auto* s = (TickSim*)p; if (slot >= s->R) return SERFSIM_E_INVAL; for (u32 v = 0; v < s->N; ++v) out[v] = s->at(slot, v).inc; return 0; }
ORC int oracle_sim_ml_state(void* p, u32 slot, u8* out) { auto* s = (TickSim*)p; if (slot >= s->R) return SERFSIM_E_INVAL; for (u32 v = 0; v < s->N; ++v) out[v] = ml_state(s->at(slot, v)); return 0; }
ORC int oracle_sim_records(void* p, u32 slot, void* out) { auto* s = (TickSim*)p; if (slot >= s->R) return SERFSIM_E_INVAL; memcpy(out, &s->at(slot, 0), (size_t)s->N * 32); return 0; }
ORC int oracle_sim_tick_trace(void* p, u32 first, u32 n, serfsim_tick_row_t* out) { auto* s = (TickSim*)p; if ((u64)first + n > s->trace.size()) return SERFSIM_E_INVAL; memcpy(out, s->trace.data() + first, (size_t)n * sizeof(serfsim_tick_row_t)); return 0; }
ORC int oracle_sim_state_hash(void* p, u64* out) { *out = ((TickSim*)p)->state_hash(); return 0; }
ORC int oracle_sim_stats(void* p, serfsim_stats_t* o) {
  auto* s = (TickSim*)p; memset(o, 0, sizeof(*o));
  o->tick = s->tick; o->members = s->N;
  for (size_t i = 0; i < s->trace.size(); ++i) {
    auto& r = s->trace[i];
    o->packets += r.packets; o->edge_updates += r.edge_updates; o->messages += r.messages; o->changed += r.changed; o->events += r.events;
    if (r.pending || r.edge_updates || r.events) o->last_active_tick = i;
  }
  if (!s->trace.empty()) o->pending = s->trace.back().pending;
  for (u32 v = 0; v < s->N; ++v) o->member_time = std::max<u64>(o->member_time, s->node[v].clock);
  for (auto& r : s->rec) o->intent_queue += (r.txj ? 1 : 0) + (r.txl ? 1 : 0);
  for (u32 sl = 0; sl < s->R; ++sl) {
    bool first = true, diff = false; u64 key0 = 0;
    for (u32 v = 0; v < s->N; ++v) {
      if (!s->node[v].up || s->subj[sl] == v) continue;
      View& r = s->at(sl, v);
      u64 key = ((u64)r.st << 32) ^ ((u64)r.inc << 8) ^ ((u64)(known(r) ? r.status : 0) << 4) ^ ml_state(r) ^ ((u64)known(r) << 63);
      if (first) { key0 = key; first = false; } else if (key != key0) diff = true;
    }
    if (diff) o->disagree_slots++;
  }
  return 0;
}

ORC void oracle_byz_stale(const void* rec32, u32 delta, u32* out /*kind, lt, key*/) { View r; memcpy(&r, rec32, 32); u8 k; byz_stale(r, delta, &k, &out[1], &out[2]); out[0] = k; }
ORC int oracle_byz_judge(const void* rec32, u32 kind, u32 val, u32 delta) { View q; memcpy(&q, rec32, 32); return byz_judge(q, (u8)kind, val, delta); }
// ---- byzantine injectors: same shapes as serfsim_set_byzantine / serfsim_anomaly_flags / serfsim_byzantine_stats ----
ORC int oracle_sim_set_byzantine(void* p, u32 n, const u32* ids, u32 delta) {
  auto* s = (TickSim*)p;
  if ((n && !ids) || s->tick != 0 || !s->events.empty()) { g_err = "bad set_byzantine"; return SERFSIM_E_INVAL; }
  if (n && s->own_count != s->N) { g_err = "byzantine injectors: single oracle instance"; return SERFSIM_E_INVAL; }
  s->byz.assign(n ? s->N : 0, 0); s->byz_n = 0;
  for (u32 i = 0; i < n; ++i) { if (ids[i] >= s->N || s->byz[ids[i]]) { g_err = "bad byzantine id"; return SERFSIM_E_INVAL; } s->byz[ids[i]] = 1; s->byz_n++; }
  s->byz_delta = delta; s->anomaly.assign(n ? s->N : 0, 0); for (auto& x : s->byz_tot) x = 0;
  return 0;
}
ORC int oracle_sim_anomaly_flags(void* p, u8* out) { auto* s = (TickSim*)p; if (!s->byz_n) return SERFSIM_E_INVAL; memcpy(out, s->anomaly.data(), s->N); return 0; }
ORC int oracle_sim_byzantine_stats(void* p, serfsim_byz_stats_t* o) {
  auto* s = (TickSim*)p; if (!s->byz_n) return SERFSIM_E_INVAL;
  o->messages = s->byz_tot[0]; o->edge_updates = s->byz_tot[1]; o->flagged = s->byz_tot[2];
  return 0;
}

// ---- user events: same shapes as serfsim_set_user_events / serfsim_user_event_* ----
ORC int oracle_sim_set_user_events(void* p, u32 n, const u32* content) {
  auto* s = (TickSim*)p;
  if (n > 8 || (n && !content) || s->tick != 0 || !s->events.empty()) { g_err = "bad set_user_events"; return SERFSIM_E_INVAL; }
  if (n && s->own_count != s->N) { g_err = "user events: single oracle instance"; return SERFSIM_E_INVAL; }
  s->ue_n = n; for (u32 e = 0; e < n; ++e) s->ue_content[e] = content[e];
  s->ue_reset();
  return 0;
}
ORC int oracle_sim_event_time(void* p, u64* out) { auto* s = (TickSim*)p; if (!s->ue_n) return SERFSIM_E_INVAL; for (u32 v = 0; v < s->N; ++v) out[v] = s->uen[v].clock; return 0; }
ORC int oracle_sim_user_event_seen(void* p, u32 e, u8* out) {
  auto* s = (TickSim*)p; if (e >= s->ue_n) return SERFSIM_E_INVAL;
  for (u32 v = 0; v < s->N; ++v) { u32 w[4]; s->ue_record(v, w); out[v] = (u8)((w[1] >> e) & 1u); }
  return 0;
}
ORC int oracle_sim_user_event_ltime(void* p, u32 e, u64* out) { auto* s = (TickSim*)p; if (e >= s->ue_n) return SERFSIM_E_INVAL; *out = s->ue_ltime[e]; return 0; }
ORC int oracle_sim_user_event_records(void* p, void* out) { auto* s = (TickSim*)p; if (!s->ue_n) return SERFSIM_E_INVAL; for (u32 v = 0; v < s->N; ++v) s->ue_record(v, (u32*)out + 4 * (size_t)v); return 0; }
ORC int oracle_sim_user_event_stats(void* p, serfsim_uevent_stats_t* o) {
  auto* s = (TickSim*)p; if (!s->ue_n) return SERFSIM_E_INVAL;
  memset(o, 0, sizeof(*o));
  o->messages = s->ue_tot[0]; o->edge_updates = s->ue_tot[1]; o->delivered = s->ue_tot[2]; o->duplicates = s->ue_tot[3]; o->too_old = s->ue_tot[4];
  for (auto& nd : s->uen) { for (u32 e = 0; e < s->ue_n; ++e) o->event_queue += nd.tx[e] ? 1 : 0; o->event_time = std::max<u64>(o->event_time, nd.clock); }
  return 0;
}

// single-node probe: a sequence of (tracked event, its ltime) arrivals through TickSim::ue_handle; reports the outcomes and the record
ORC void oracle_ue_handle_seq(u32 n_events, const u32* content, const u32* ltime_tab, u32 limit, u32 n, const u32* ev, int* outcomes, u32* record) {
  TickSim s; s.N = 1; s.R = 1; s.ue_n = n_events;
  for (u32 e = 0; e < n_events; ++e) { s.ue_content[e] = content[e]; s.ue_ltime[e] = ltime_tab[e]; }
  s.uen.assign(1, UeNodeB{});
  for (u32 i = 0; i < n; ++i) { outcomes[i] = s.ue_handle(s.uen[0], ev[i], ltime_tab[ev[i]]); if (outcomes[i] == 0) s.uen[0].tx[ev[i]] = (u8)limit; }
  s.ue_record(0, record);
}

// =====================================================================================
// Part C — FaithfulSim: N literal serf nodes (Part A), each with its own full member table and a
// multi-entry TransmitLimitedQueue, driven tick by tick with the SAME peer selection as Part B.
// Used only by tests to check that the packed-record tick model (two queue entries per view, inbox
// reduction semantics) agrees with a literal multi-node execution on serf-only scenarios (join /
// force-leave operations; the memberlist layer is not part of this model).
// =====================================================================================
struct FaithfulSim {
  u32 N, fanout, retransmit_mult; u64 seed; u32 tick = 0;
  std::vector<RefNode> nodes;
  std::vector<u64> row_ptr; std::vector<u32> col;
  struct Op { u32 tick, op, node; u64 subject; };
  std::vector<Op> ops;
  struct M { u32 dst, src; u8 ty; u64 ltime, id; bool prune; };
  std::vector<M> inflight;
  std::vector<u64> subjects;      // processing order of subjects at a receiver (slot order of Part B)

  FaithfulSim(u32 n, u32 f, u32 rm, u64 sd, u64 init_st, u64 init_clock) : N(n), fanout(f), retransmit_mult(rm), seed(sd) {
    nodes.reserve(n);
    for (u32 v = 0; v < n; ++v) {
      nodes.emplace_back((u64)v);
      RefNode& nd = nodes.back();
      nd.retransmit_mult = rm;
      nd.clock.v = init_clock;
      for (u32 w = 0; w < n; ++w) nd.states[w] = MemberStateA{ST_ALIVE, init_st, false, 0};
    }
  }
  u32 targets(u32 v, u32 t, u32* out) const {      // identical to TickSim::gossip_targets
    u64 r0 = row_ptr[v]; u32 deg = (u32)(row_ptr[v + 1] - r0), nt = 0;
    const u32 m = std::min(fanout, deg);
    if (!m) return 0;
    u32 w[4];
    philox4x32_10(t, v, 0, DOMAIN_GOSSIP, (u32)seed, (u32)(seed >> 32), w);
    u32 chosen[8], nc = 0;
    for (u32 k = 0; k < m; ++k) {
      const u32 x = w[(k >> 1) & 3];
      u32 j = (((k & 1) ? (x >> 16) : (x & 0xffffu)) * (deg - k)) >> 16;
      for (u32 i = 0; i < nc; ++i) if (j >= chosen[i]) ++j;
      u32 pos = nc;
      while (pos > 0 && chosen[pos - 1] > j) { chosen[pos] = chosen[pos - 1]; --pos; }
      chosen[pos] = j; ++nc;
      const u32 c = col[r0 + j];
      if (c != v) out[nt++] = c;
    }
    return nt;
  }
  int order_of(u64 id) const { for (size_t i = 0; i < subjects.size(); ++i) if (subjects[i] == id) return (int)i; return (int)subjects.size(); }
  void step_one() {
    const u32 t = tick;
    // Phase R: canonical order per receiver: subjects in slot order, leaves ascending, then joins ascending
    std::stable_sort(inflight.begin(), inflight.end(), [&](const M& a, const M& b) {
      if (a.dst != b.dst) return a.dst < b.dst;
      const int oa = order_of(a.id), ob = order_of(b.id);
      if (oa != ob) return oa < ob;
      if (a.ty != b.ty) return a.ty < b.ty;            // TY_LEAVE (1) before TY_JOIN (2)
      if (a.ltime != b.ltime) return a.ltime < b.ltime;
      if (a.prune != b.prune) return a.prune > b.prune;   // equal Lamport time: the pruning intent first (Part B's canonical order)
      return a.src < b.src;
    });
    // Per node, per subject in slot order: Phase R (that subject's messages, leaves then joins), the refutation task,
    // then Phase E (a host operation of this tick that concerns that subject) — the interleaving Part B uses.
    size_t i = 0;
    for (u32 d = 0; d < N; ++d) {
      RefNode& nd = nodes[d];
      const Op* op = nullptr;
      for (auto& o : ops) if (o.tick == t && o.node == d) { op = &o; break; }
      const size_t end0 = i;
      size_t end = end0;
      while (end < inflight.size() && inflight[end].dst == d) ++end;
      if (end == end0 && !op) continue;
      for (size_t sl = 0; sl <= subjects.size(); ++sl) {           // the last round takes untracked subjects (none in the tests)
        while (i < end && order_of(inflight[i].id) == (int)sl) {
          const M& m = inflight[i++];
          const bool rb = (m.ty == TY_JOIN) ? nd.handle_node_join_intent(m.ltime, m.id) : nd.handle_node_leave_intent(m.ltime, m.id, m.prune);
          if (rb) nd.queue(m.ty, m.ltime, m.id, m.prune, false);         // serf/delegate.rs:294-300: the raw message is re-queued, prune flag included
        }
        nd.run_detached();                                               // the refutation task of this subject, if any
        if (op && sl < subjects.size()) {
          if (op->op == SERFSIM_OP_JOIN && subjects[sl] == d) nd.api_join();
          else if ((op->op == SERFSIM_OP_FORCE_LEAVE || op->op == SERFSIM_OP_FORCE_LEAVE_PRUNE) && subjects[sl] == op->subject) { nd.api_force_leave(op->subject, op->op == SERFSIM_OP_FORCE_LEAVE_PRUNE); nd.run_detached(); }
        }
      }
      i = end;
    }
    inflight.clear();
    // Phase S: every node with queued broadcasts sends one packet (all entries that fit) to each gossip peer
    for (u32 v = 0; v < N; ++v) {
      RefNode& nd = nodes[v];
      if (nd.broadcasts.empty()) continue;
      u32 tg[8]; const u32 nt = targets(v, t, tg);
      for (u32 k = 0; k < nt && !nd.broadcasts.empty(); ++k) {
        u8 ty[64], pr[64]; u64 lt[64], id[64];
        const u32 n = nd.get_broadcasts(1u << 20, 2, ty, lt, id, 64, pr);
        for (u32 q = 0; q < n && q < 64; ++q) inflight.push_back(M{tg[k], v, ty[q], lt[q], id[q], pr[q] != 0});
      }
    }
    ++tick;
  }
};

ORC void* faithful_new(u32 n, u32 fanout, u32 retransmit_mult, u64 seed, u64 init_st, u64 init_clock) { return new FaithfulSim(n, fanout, retransmit_mult, seed, init_st, init_clock); }
ORC void faithful_free(void* p) { delete (FaithfulSim*)p; }
ORC void faithful_set_topology(void* p, const u64* row_ptr, const u32* col) { auto* s = (FaithfulSim*)p; s->row_ptr.assign(row_ptr, row_ptr + s->N + 1); s->col.assign(col, col + row_ptr[s->N]); }
ORC void faithful_set_subjects(void* p, const u64* subj, u32 n) { ((FaithfulSim*)p)->subjects.assign(subj, subj + n); }
ORC void faithful_inject(void* p, u32 tick, u32 op, u32 node, u64 subject) { ((FaithfulSim*)p)->ops.push_back(FaithfulSim::Op{tick, op, node, subject}); }
ORC void faithful_step(void* p, u32 n) { for (u32 i = 0; i < n; ++i) ((FaithfulSim*)p)->step_one(); }
ORC int faithful_view(void* p, u32 node, u64 subject, u8* status, u64* st) { auto& nd = ((FaithfulSim*)p)->nodes[node]; auto it = nd.states.find(subject); if (it == nd.states.end()) return 0; *status = it->second.status; *st = it->second.status_time; return 1; }
ORC u64 faithful_clock(void* p, u32 node) { return ((FaithfulSim*)p)->nodes[node].clock.time(); }
ORC u32 faithful_queue_len(void* p, u32 node) { return (u32)((FaithfulSim*)p)->nodes[node].broadcasts.size(); }
ORC u32 faithful_inflight(void* p) { return (u32)((FaithfulSim*)p)->inflight.size(); }

// Sharded runs of the tick oracle (tests of the N > 1 host logic): an instance owns a contiguous id range, processes only
// those nodes, and hands the messages for foreign nodes to the caller, who delivers them to the owning instance.
ORC int oracle_sim_set_ownership(void* p, u32 first, u32 count) {
  auto* s = (TickSim*)p;
  if (s->tick != 0 || (u64)first + count > s->N || count == 0) return SERFSIM_E_INVAL;
  s->own_first = first; s->own_count = count; return 0;
}
ORC u32 oracle_sim_export(void* p, void* out, u32 cap) {          // 16-byte entries {dst, src, val, slot u8, kind u8, pad}
  auto* s = (TickSim*)p; u32 n = 0;
  for (auto& e : s->exported) for (auto& m : e) { if (n < cap) memcpy((char*)out + (size_t)n * sizeof(Msg), &m, sizeof(Msg)); ++n; }
  return n;
}
ORC int oracle_sim_import(void* p, const void* in, u32 n) {
  auto* s = (TickSim*)p; const u32 T = (u32)s->threads;
  if (s->mail.size() != (size_t)T * T) s->mail.assign((size_t)T * T, {});
  for (u32 i = 0; i < n; ++i) {
    Msg m; memcpy(&m, (const char*)in + (size_t)i * sizeof(Msg), sizeof(Msg));
    if (m.dst - s->own_first >= s->own_count) return SERFSIM_E_INVAL;
    s->mail[(size_t)0 * T + s->owner_of(m.dst)].push_back(m);
  }
  return 0;
}
ORC u32 oracle_msg_size(void) { return (u32)sizeof(Msg); }

// Pin worker c of the tick loop to logical CPU cpus[c % n] (n = 0: no pinning).  Timing aid for bench.py's CPU arm: unpinned
// workers migrate between NUMA nodes and the same run varied 5x between two boxes.  Results do not depend on it.
ORC int oracle_sim_set_affinity(void* p, const int* cpus, int n) {
  auto* s = (TickSim*)p;
  if (n < 0 || (n && !cpus)) return SERFSIM_E_INVAL;
  s->pin.assign(cpus, cpus + n);
  return 0;
}
// Number of host threads the tick loop uses (results are independent of it).  Takes effect at the next reset.
ORC int oracle_sim_set_threads(void* p, int n) {
  auto* s = (TickSim*)p;
  if (n < 1 || s->tick != 0) return SERFSIM_E_INVAL;
  s->threads = n; s->chunk = (s->N + n - 1) / n; s->mail.clear(); s->mail_next.clear(); s->scratch.clear();
  return 0;
}
