// Host-compiled check of the user-event rules the CUDA kernel runs (serf_b200/csrc/uevent.cuh).
//
// uevent.cuh is __host__ __device__ code without memory traffic; uevent_kernel.cu wraps it in loads, stores and one
// RED.OR per target.  This file wraps the SAME functions in a sequential loop with the same data layout (two inbox
// planes by tick parity, 16-byte records, the stamped-ltime table) so that tests/test_uevent_rules.py can compare the
// packed-mask model with the oracle's literal ring-buffer model on a machine without a GPU.  Test infrastructure only:
// nothing in the product calls it.
#include <cstdint>
#include <cstring>
#include <vector>
struct uint4 { unsigned x, y, z, w; };
#define __host__
#define __device__
#include "../../serf_b200/csrc/uevent.cuh"

using namespace sfs;

extern "C" __attribute__((visibility("default")))
int uecheck_run(u32 N, const u64* row_ptr, const u32* col, u32 fanout, u64 seed, u32 limit, u32 R,
                u32 n_events, const u32* content,
                u32 n_ops, const u32* op_tick, const u32* op_kind, const u32* op_node, const u32* op_slot,
                u32 n_ticks,
                u32* out_records /*[N][4]*/, u64* out_rows /*[n_ticks][5]: edge_updates, messages, changed, pending, hash*/,
                u64* out_totals /*[5]*/, u32* out_ltime /*[8]*/) {
  if (n_events > MAX_UEVENTS || fanout > MAX_FANOUT) return -1;
  UeTable tb{};
  tb.n = n_events;
  for (u32 e = 0; e < n_events; ++e) tb.content[e] = content[e];
  std::vector<uint4> state(N, uint4{UE_INIT_CLOCK, 0, 0, 0});
  std::vector<u32> inbox[2] = {std::vector<u32>(N, 0), std::vector<u32>(N, 0)};
  std::vector<u8> up(N, 1);
  u32 ltime[MAX_UEVENTS] = {0};
  u64 totals[5] = {0, 0, 0, 0, 0};
  for (u32 t = 0; t < n_ticks; ++t) {
    std::vector<u32>& inbox_rd = inbox[(t & 1) ^ 1];
    std::vector<u32>& inbox_wr = inbox[t & 1];
    u64 row[5] = {0, 0, 0, 0, 0};
    std::vector<u8> up_next = up;
    for (u32 v = 0; v < N; ++v) {
      const u32 arrived = inbox_rd[v];
      if (arrived) inbox_rd[v] = 0;
      u32 op = 0, opslot = 0;
      for (u32 i = 0; i < n_ops; ++i) if (op_tick[i] == t && op_node[i] == v) { op = op_kind[i]; opslot = op_slot[i]; break; }
      const uint4 w0 = state[v];
      UeRec r;
      ue_unpack(w0, r);
      const bool up_r = up[v] != 0;
      bool up_s = up_r;
      if (op == OP_FAIL) up_s = false;
      if (op == OP_REJOIN) up_s = true;
      up_next[v] = up_s;
      UeCounts c{};
      bool stamped = false;
      const u32 L = ue_receive_and_originate(r, arrived, up_r, op, opslot, ltime, tb, limit, c, stamped);
      if (stamped) ltime[opslot] = L;
      row[2] += c.delivered;
      if (up_s && ue_queued(r, tb.n)) {
        const u32 row0 = (u32)row_ptr[v], deg = (u32)(row_ptr[v + 1] - row_ptr[v]);
        u32 tg[MAX_FANOUT];
        const u32 nt = ue_pick_targets(t, v, row0, deg, fanout, (u32)seed, (u32)(seed >> 32), col, tg);
        u32 bits[MAX_FANOUT];
        c.messages += ue_plan_send<(int)MAX_FANOUT>(r, tb.n, nt, bits);
        for (u32 k = 0; k < nt; ++k) if (bits[k]) { inbox_wr[tg[k]] |= bits[k]; c.edges++; }
      }
      if (up_s) c.pending += ue_queued(r, tb.n);
      const uint4 w1 = ue_pack(r);
      state[v] = w1;
      row[0] += c.edges; row[1] += c.messages; row[3] += c.pending;
      row[4] += ue_hash((u64)(R + 1) * N + v, w1);
      totals[0] += c.messages; totals[1] += c.edges; totals[2] += c.delivered; totals[3] += c.duplicates; totals[4] += c.too_old;
    }
    up = up_next;
    memcpy(out_rows + (size_t)t * 5, row, sizeof(row));
  }
  for (u32 v = 0; v < N; ++v) { out_records[4 * (size_t)v] = state[v].x; out_records[4 * (size_t)v + 1] = state[v].y; out_records[4 * (size_t)v + 2] = state[v].z; out_records[4 * (size_t)v + 3] = state[v].w; }
  memcpy(out_totals, totals, sizeof(totals));
  memcpy(out_ltime, ltime, sizeof(ltime));
  return 0;
}

// single-node probe: run a sequence of (event, ltime) arrivals through ue_handle and report outcomes + the packed record
extern "C" __attribute__((visibility("default")))
void uecheck_handle_seq(u32 n_events, const u32* content, const u32* ltime_tab, u32 limit, u32 n, const u32* ev, int* outcomes, u32* record /*[4]*/) {
  UeTable tb{};
  tb.n = n_events;
  for (u32 e = 0; e < n_events; ++e) tb.content[e] = content[e];
  UeRec r{};
  r.clock = UE_INIT_CLOCK;
  for (u32 i = 0; i < n; ++i) outcomes[i] = ue_handle(r, ev[i], ltime_tab[ev[i]], ltime_tab, tb, limit, true);
  const uint4 w = ue_pack(r);
  record[0] = w.x; record[1] = w.y; record[2] = w.z; record[3] = w.w;
}
