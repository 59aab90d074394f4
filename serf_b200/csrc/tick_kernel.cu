// tick_kernel.cu — the fused gossip tick for sm_100a.
//
// One launch = one gossip tick of every virtual node of this shard:
//   Phase R  receive: fold the reduced inbox of the previous tick into the node's views
//            (SerfDelegate::notify_message → handle_node_{join,leave}_intent, serf/delegate.rs:157-315,
//             serf/base.rs:1338-1373, 1442-1572; LamportClock::witness, types/clock.rs:155-172;
//             memberlist alive/suspect/dead merge [external]);
//   Phase E  host operations scheduled for this tick (Serf::join/leave/remove_failed_node, serf/api.rs);
//   Phase T  suspicion-timer expiry and the SWIM probe [external];
//   Phase S  gossip: pick `fanout` distinct peers with the counter RNG and reduce every queued
//            entry into their inbox with RED.MAX (SerfDelegate::broadcast_messages + the
//            TransmitLimitedQueue budget, serf/delegate.rs:317-384, serf/base.rs:179-190).
// Sends of tick t land in inbox parity t&1 and are consumed by Phase R of tick t+1, so a launch
// never reads what it writes: bulk-synchronous, order-independent, bit-reproducible.
//
// Memory behaviour (HBM-bound integer work, no tensor cores): one thread per node streams its
// R 32-byte records with 128-bit loads (a warp covers 1 KB contiguous per plane), the three
// inbox planes with coalesced 32-bit loads, and scatters 32-bit RED.MAX to random peers; the
// inbox planes are the only randomly addressed data and are sized to stay L2-resident.
#include <cooperative_groups.h>

#include "tick_kernel.cuh"

namespace sfs {

namespace {

constexpr int BLOCK = 256;

__device__ __forceinline__ uint4 ld_rec(const uint4* p) { return __ldcg(p); }        // L2-only: records are streamed once per tick
__device__ __forceinline__ void st_rec(uint4* p, const uint4& v) { __stcg(p, v); }

struct Counters {
  u32 packets, edges, msgs, changed, pending, events, suspects, kL, kJ, kM;
  u64 hash;
};

__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deliver one entry to `dst` (global id): local → RED.MAX into this shard's inbox.
__device__ __forceinline__ void deliver(const TickParams& p, u32 dst, u32 kind, u32 s, u32 val1) {
  u32 dl = dst - p.first;
  if (dl < p.n_local) {
    atomicMax(p.inbox_wr + ((size_t)(kind * p.R + s)) * p.n_local + dl, val1);
  } else {
    // cross-shard: append (dst_local, kind, slot, value) to the peer's receive window over NVLink
    u32 shard = dst / p.shard_size;
    u32 dloc = dst - shard * p.shard_size;
    u32 pos = atomicAdd(p.win_count[shard] + p.rank, 1u);
    if (pos < p.win_cap) {
      u64 e = ((u64)val1 << 32) | ((u64)s << 28) | ((u64)kind << 26) | dloc;
      p.win_data[shard][(size_t)p.rank * p.win_cap + pos] = e;
    } else {
      *p.overflow = 2;
    }
  }
}

template <bool TRACE>
__device__ __forceinline__ void process_node(const TickParams& p, u32 vl, bool kL, bool kJ, bool kM, Counters& c) {
  const u32 v = p.first + vl;
  const u32 t = p.tick;
  const u32 limit = p.rules.limit;
  const u64 ns = p.node_state[vl];
  u32 clock = (u32)ns;
  const bool up_r = (ns & NS_UP) != 0;
  u32 sstate = (u32)(ns >> 40) & 3;

  // host operation for this node (at most one per tick; the mark kernel set NS_EV)
  u32 op = 0, op_slot = 0;
  if (ns & NS_EV) {
    for (u32 e = p.ev_begin; e < p.ev_end; ++e)
      if (p.ev_node[e] == v) { op = p.ev_op[e]; op_slot = p.ev_slot[e]; break; }
    c.events++;
  }
  bool up_s = up_r;
  if (op == OP_FAIL) up_s = false;
  if (op == OP_REJOIN) up_s = true;

  const u32 row0 = p.row_ptr[vl];
  const u32 deg = p.row_ptr[vl + 1] - row0;

  // SWIM probe target of this round (only matters while some tracked subject is down)
  bool have_probe = false;
  u32 ptarget = 0;
  if (up_s && p.probe_every && p.down_mask && ((t + v) % p.probe_every) == 0 && deg) {
    u32 w[4];
    philox4x32_10(t, v, 0, DOMAIN_PROBE, p.seed_lo, p.seed_hi, w);
    ptarget = __ldg(p.col + row0 + mulhi32(w[0], deg));
    have_probe = true;
  }

  u32 targets[MAX_FANOUT];
  u32 nt = 0;
  bool have_targets = false;
  u32 max_tx = 0;

  for (u32 s = 0; s < p.R; ++s) {
    const size_t idx = (size_t)s * p.n_local + vl;
    const uint4 a0 = ld_rec(p.rec + 2 * idx), b0 = ld_rec(p.rec + 2 * idx + 1);
    u32 mL = 0, mJ = 0, mM = 0;
    if (kL) { u32* q = p.inbox_rd + ((size_t)(KIND_LEAVE * p.R + s)) * p.n_local + vl; mL = __ldcg(q); if (mL) __stcg(q, 0u); }
    if (kJ) { u32* q = p.inbox_rd + ((size_t)(KIND_JOIN * p.R + s)) * p.n_local + vl; mJ = __ldcg(q); if (mJ) __stcg(q, 0u); }
    if (kM) { u32* q = p.inbox_rd + ((size_t)(KIND_ML * p.R + s)) * p.n_local + vl; mM = __ldcg(q); if (mM) __stcg(q, 0u); }
    Rec r;
    unpack(a0, b0, r);
    const bool self = (p.subj[s] == v);

    // ---------------- Phase R ----------------
    if (up_r && (mL | mJ | mM)) {
      if (mM) {
        const u32 key = mM - 1, inc = key >> 6, kind = (key >> 4) & 3, fromb = key & 15;
        if (kind == ML_ALIVE) ml_alive(r, inc, self, limit);
        else if (kind == ML_SUSPECT) ml_suspect(r, inc, fromb, t, self, p.rules);
        else ml_dead(r, inc, kind == ML_LEFT, t, self, limit);
      }
      bool refute = false;
      if (mL) { const u32 lt = mL - 1; witness(clock, lt); leave_intent(r, lt, self, sstate, refute, limit); }
      if (mJ) { const u32 lt = mJ - 1; witness(clock, lt); join_intent(r, lt, limit); }
      if (refute) {                                       // serf/base.rs:1470-1480 → broadcast_join(clock.time()), :381-397
        const u32 T = clock; witness(clock, T);
        join_intent(r, T, limit);
        r.qjoin = T; r.txj = limit;
      }
      uint4 a1, b1;
      pack(r, a1, b1);
      if (a1.x != a0.x || a1.y != a0.y || a1.z != a0.z || a1.w != a0.w || b1.x != b0.x || b1.y != b0.y || b1.z != b0.z || b1.w != b0.w)
        c.changed++;
    }
    // ---------------- Phase E ----------------
    if (op) {
      if (op == OP_REJOIN && self && !up_r) {
        r.inc += 1; r.mlstate = ML_ALIVE; r.qfrom = 0; r.txm = limit; r.deadline = 0; r.mask = 0;
        sstate = SS_ALIVE;
        node_join(r);
      }
      if (((op == OP_JOIN && up_r) || (op == OP_REJOIN && !up_r)) && self) {     // serf/api.rs:339-342 → serf/base.rs:381-397
        const u32 T = clock; witness(clock, T);
        join_intent(r, T, limit);
        r.qjoin = T; r.txj = limit;
      }
      if (op == OP_LEAVE && up_r && self && sstate == SS_ALIVE) {                // serf/api.rs:422-449
        sstate = SS_LEAVING;
        const u32 T = clock; clock += 1;
        bool rf = false;
        leave_intent(r, T, true, sstate, rf, limit);
        r.qleave = T; r.txl = limit;
      }
      if (op == OP_FORCE_LEAVE && up_r && op_slot == s) {                        // serf/base.rs:454-480
        const u32 T = clock; witness(clock, T);
        bool rf = false;
        leave_intent(r, T, self, sstate, rf, limit);
        r.qleave = T; r.txl = limit;
        if (rf) { const u32 T2 = clock; witness(clock, T2); join_intent(r, T2, limit); r.qjoin = T2; r.txj = limit; }
      }
    }
    if (up_s) {
      // ---------------- Phase T ----------------
      if (r.mlstate == ML_SUSPECT && r.deadline != 0 && t >= r.deadline) ml_dead(r, r.inc, false, t, false, limit);
      if (have_probe && !self && ptarget == p.subj[s] && ((p.down_mask >> s) & 1)) {
        if (r.mlstate == ML_ALIVE || r.mlstate == ML_SUSPECT) {
          if (r.mlstate == ML_ALIVE) c.suspects++;
          ml_suspect(r, r.inc, from_bucket(v), t, false, p.rules);
        }
      }
      // ---------------- Phase S ----------------
      if (r.txl | r.txj | r.txm) {
        if (!have_targets) {
          // kRandomNodes: up to 3·deg draws for `fanout` distinct peers other than ourselves
          u32 w[4] = {0, 0, 0, 0};
          for (u32 i = 0; i < 3 * deg && nt < p.fanout; ++i) {
            if ((i & 3) == 0) philox4x32_10(t, v, i >> 2, DOMAIN_GOSSIP, p.seed_lo, p.seed_hi, w);
            const u32 cnd = __ldg(p.col + row0 + mulhi32(w[i & 3], deg));
            bool skip = (cnd == v);
#pragma unroll
            for (u32 j = 0; j < MAX_FANOUT; ++j) skip |= (j < nt && targets[j] == cnd);
            if (!skip) {
#pragma unroll
              for (u32 j = 0; j < MAX_FANOUT; ++j) if (j == nt) targets[j] = cnd;
              ++nt;
            }
          }
          have_targets = true;
        }
        const u32 key1 = ml_key(r) + 1;
#pragma unroll
        for (u32 k = 0; k < MAX_FANOUT; ++k) {
          if (k < nt) {
            u32 cnt = 0;
            if (r.txl > k) { deliver(p, targets[k], KIND_LEAVE, s, r.qleave + 1); ++cnt; c.kL++; }
            if (r.txj > k) { deliver(p, targets[k], KIND_JOIN, s, r.qjoin + 1); ++cnt; c.kJ++; }
            if (r.txm > k) { deliver(p, targets[k], KIND_ML, s, key1); ++cnt; c.kM++; }
            if (cnt) { c.edges++; c.msgs += cnt; }
          }
        }
        max_tx = max(max_tx, max(r.txl, max(r.txj, r.txm)));
        r.txl -= min(r.txl, nt); r.txj -= min(r.txj, nt); r.txm -= min(r.txm, nt);
      }
      // Serf::leave: our own leave intent is out → memberlist.leave() → dead{node == from}  (serf/api.rs:451-476)
      if (self && sstate == SS_LEAVING && r.txl == 0 && r.mlstate == ML_ALIVE) {
        r.mlstate = ML_LEFT; r.qfrom = 0; r.txm = limit; sstate = SS_LEFT;
      }
      const bool pend = (r.txl | r.txj | r.txm) || r.mlstate == ML_SUSPECT ||
                        (p.probe_every && ((p.down_mask >> s) & 1) && !self && r.mlstate == ML_ALIVE);
      c.pending += pend ? 1 : 0;
    }
    uint4 a2, b2;
    pack(r, a2, b2);
    if (a2.x != a0.x || a2.y != a0.y || a2.z != a0.z || a2.w != a0.w) st_rec(p.rec + 2 * idx, a2);
    if (b2.x != b0.x || b2.y != b0.y || b2.z != b0.z || b2.w != b0.w) st_rec(p.rec + 2 * idx + 1, b2);
    if (TRACE) c.hash += rec_hash((u64)s * p.n_global + v, a2, b2);
    if (r.inc >= INC_LIMIT) *p.overflow = 1;
  }
  const u64 ns2 = (u64)clock | (up_s ? NS_UP : 0) | ((u64)sstate << 40);
  if (ns2 != ns) p.node_state[vl] = ns2;
  if (TRACE) c.hash += node_hash((u64)p.R * p.n_global + v, ns2);
  if (clock >= LTIME_LIMIT) *p.overflow = 1;
  c.packets += min(nt, max_tx);
}

template <bool TRACE>
__global__ void __launch_bounds__(BLOCK) tick_kernel(const __grid_constant__ TickParams p) {
  Counters c = {};
  const bool kL = p.kinds_prev[KIND_LEAVE] != 0, kJ = p.kinds_prev[KIND_JOIN] != 0, kM = p.kinds_prev[KIND_ML] != 0;
  for (u32 base = blockIdx.x * BLOCK; base < p.n_local; base += gridDim.x * BLOCK) {
    const u32 vl = base + threadIdx.x;
    if (vl < p.n_local) process_node<TRACE>(p, vl, kL, kJ, kM, c);
  }
  // block reduction → one atomic per counter per CTA
  __shared__ u64 red[12][BLOCK / 32];
  u64 vals[12] = {c.packets, c.edges, c.msgs, c.changed, c.pending, c.events, c.suspects, c.hash, c.kL, c.kJ, c.kM, 0};
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 11; ++i) {
    u64 s = (i == 7) ? warp_sum64(vals[i]) : (u64)warp_sum((u32)vals[i]);
    if (lane == 0) red[i][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 11) {
    u64 s = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) s += red[threadIdx.x][w];
    if (s) {
      if (threadIdx.x < 8) atomicAdd((unsigned long long*)(p.row + threadIdx.x), (unsigned long long)s);
      else atomicAdd(p.kinds_cur + (threadIdx.x - 8), (u32)min(s, (u64)0xffffffffu));
    }
  }
}

// Fold the cross-shard window (filled by the peers during their tick kernel) into the inbox.
__global__ void __launch_bounds__(BLOCK) drain_kernel(const __grid_constant__ DrainParams p) {
  for (u32 src = 0; src < p.world; ++src) {
    if (src == p.rank) continue;
    const u32 n = min(p.win_count[src], p.win_cap);
    const u64* w = p.win_data + (size_t)src * p.win_cap;
    for (u32 i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += gridDim.x * BLOCK) {
      const u64 e = __ldcg(w + i);
      const u32 val1 = (u32)(e >> 32), s = (u32)(e >> 28) & 15, kind = (u32)(e >> 26) & 3, dl = (u32)e & ((1u << 26) - 1);
      if (dl < p.n_local && s < p.R && kind < 3) atomicMax(p.inbox_wr + ((size_t)(kind * p.R + s)) * p.n_local + dl, val1);
      else *p.overflow = 3;
    }
  }
}

__global__ void init_state_kernel(uint4* rec, u64* node_state, u32 n_local, u32 R, u32 init_st, u32 init_clock) {
  const u32 vl = blockIdx.x * blockDim.x + threadIdx.x;
  if (vl >= n_local) return;
  Rec r = {};
  r.st = init_st; r.inc = 1; r.status = ST_ALIVE; r.mlstate = ML_ALIVE; r.flags = 1;
  uint4 a, b;
  pack(r, a, b);
  for (u32 s = 0; s < R; ++s) {
    const size_t idx = (size_t)s * n_local + vl;
    rec[2 * idx] = a; rec[2 * idx + 1] = b;
  }
  node_state[vl] = (u64)init_clock | NS_UP;
}

__global__ void mark_events_kernel(u64* node_state, const u32* ev_node, u32 ev_begin, u32 ev_end, u32 first, u32 n_local) {
  const u32 e = ev_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ev_end) return;
  const u32 vl = ev_node[e] - first;
  if (vl < n_local) node_state[vl] |= NS_EV;       // one op per (node, tick): no two threads touch the same word
}

__global__ void extract_kernel(const uint4* rec, const u64* node_state, u32 n_local, u32 slot, int what, void* out) {
  const u32 vl = blockIdx.x * blockDim.x + threadIdx.x;
  if (vl >= n_local) return;
  if (what == EXTRACT_CLOCK) { ((u64*)out)[vl] = node_state[vl] & 0xffffffffull; return; }
  const size_t idx = (size_t)slot * n_local + vl;
  Rec r;
  unpack(rec[2 * idx], rec[2 * idx + 1], r);
  const bool known = r.flags & 1;
  switch (what) {
    case EXTRACT_STATUS: ((u8*)out)[vl] = known ? (u8)r.status : (u8)ST_NONE; break;
    case EXTRACT_STATUS_LTIME: ((u64*)out)[vl] = known ? r.st : 0; break;
    case EXTRACT_INC: ((u32*)out)[vl] = r.inc; break;
    case EXTRACT_ML: ((u8*)out)[vl] = (u8)r.mlstate; break;
  }
}

__global__ void __launch_bounds__(BLOCK) state_hash_kernel(const uint4* rec, const u64* node_state, u32 n_local, u32 first, u32 n_global, u32 R, u64* out) {
  u64 h = 0;
  for (u32 vl = blockIdx.x * BLOCK + threadIdx.x; vl < n_local; vl += gridDim.x * BLOCK) {
    for (u32 s = 0; s < R; ++s) {
      const size_t idx = (size_t)s * n_local + vl;
      h += rec_hash((u64)s * n_global + first + vl, rec[2 * idx], rec[2 * idx + 1]);
    }
    h += node_hash((u64)R * n_global + first + vl, node_state[vl]);
  }
  h = warp_sum64(h);
  if ((threadIdx.x & 31) == 0 && h) atomicAdd((unsigned long long*)out, (unsigned long long)h);
}

// out[0] = max clock, out[1] = queued intents, out[2+2s] = min key, out[3+2s] = max key of slot s over
// up nodes other than the subject (agreement check for Stats / convergence studies).
__global__ void __launch_bounds__(BLOCK) summary_kernel(const uint4* rec, const u64* node_state, u32 n_local, u32 first, u32 R, const u32* subj, u64* out) {
  u64 maxclock = 0, queued = 0;
  for (u32 s = 0; s < R; ++s) {
    u64 kmin = ~0ull, kmax = 0;
    const u32 sid = subj[s];
    for (u32 vl = blockIdx.x * BLOCK + threadIdx.x; vl < n_local; vl += gridDim.x * BLOCK) {
      const u64 ns = node_state[vl];
      if (s == 0) maxclock = max(maxclock, (u64)(ns & 0xffffffffull));
      const size_t idx = (size_t)s * n_local + vl;
      Rec r;
      unpack(rec[2 * idx], rec[2 * idx + 1], r);
      queued += (r.txj ? 1 : 0) + (r.txl ? 1 : 0);
      if ((ns & NS_UP) && sid != first + vl) {
        const bool known = r.flags & 1;
        const u64 key = (((u64)r.st << 32) ^ ((u64)r.inc << 8) ^ ((u64)(known ? r.status : 0) << 4) ^ r.mlstate ^ ((u64)known << 63));
        kmin = min(kmin, key); kmax = max(kmax, key);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin((unsigned long long*)(out + 2 + 2 * s), (unsigned long long)kmin);
      atomicMax((unsigned long long*)(out + 3 + 2 * s), (unsigned long long)kmax);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxclock = max(maxclock, __shfl_xor_sync(0xffffffffu, maxclock, o));
  queued = warp_sum64(queued);
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)out, (unsigned long long)maxclock);
    if (queued) atomicAdd((unsigned long long*)(out + 1), (unsigned long long)queued);
  }
}

}  // namespace

int tick_grid_size(u32 n_local) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int tiles = (int)((n_local + BLOCK - 1) / BLOCK);
  const int cap = sms * 8;                       // persistent: a multiple of the SM count (148 × 8 resident CTAs of 256)
  return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

void launch_tick(const TickParams& p, bool trace, int grid, cudaStream_t st) {
  if (trace) tick_kernel<true><<<grid, BLOCK, 0, st>>>(p);
  else tick_kernel<false><<<grid, BLOCK, 0, st>>>(p);
}
void launch_drain(const DrainParams& p, cudaStream_t st) { drain_kernel<<<148 * 4, BLOCK, 0, st>>>(p); }
void launch_init_state(uint4* rec, u64* node_state, u32 n_local, u32 R, u32 init_st, u32 init_clock, cudaStream_t st) {
  init_state_kernel<<<(n_local + 255) / 256, 256, 0, st>>>(rec, node_state, n_local, R, init_st, init_clock);
}
void launch_mark_events(u64* node_state, const u32* ev_node, u32 ev_begin, u32 ev_end, u32 first, u32 n_local, cudaStream_t st) {
  const u32 n = ev_end - ev_begin;
  if (!n) return;
  mark_events_kernel<<<(n + 127) / 128, 128, 0, st>>>(node_state, ev_node, ev_begin, ev_end, first, n_local);
}
void launch_extract(const uint4* rec, const u64* node_state, u32 n_local, u32 slot, int what, void* out, cudaStream_t st) {
  extract_kernel<<<(n_local + 255) / 256, 256, 0, st>>>(rec, node_state, n_local, slot, what, out);
}
void launch_state_hash(const uint4* rec, const u64* node_state, u32 n_local, u32 first, u32 n_global, u32 R, u64* out, cudaStream_t st) {
  state_hash_kernel<<<148 * 4, BLOCK, 0, st>>>(rec, node_state, n_local, first, n_global, R, out);
}
void launch_summary(const uint4* rec, const u64* node_state, u32 n_local, u32 first, u32 R, const u32* subj_dev, u64* out, cudaStream_t st) {
  summary_kernel<<<148 * 4, BLOCK, 0, st>>>(rec, node_state, n_local, first, R, subj_dev, out);
}

}  // namespace sfs
