// byz.cuh — byzantine stale-record injectors (BASELINE configs[4]).
//
// There is NO reference semantics for this configuration: serf silently ignores a stale intent (handle_node_join_intent
// / handle_node_leave_intent return false on `ltime <= status_time`, serf/base.rs:1346-1348, 1464-1466) and memberlist
// ignores a stale incarnation.  The model below is defined by this repository's oracle (oracle/serf_oracle.cpp,
// "byzantine" block) and implemented bit-exactly here; the rules are host/device code so tests can run them on the CPU.
//
// A byzantine node behaves honestly in every respect, and additionally, every tick it is up, re-injects for every
// subject it knows a STALE copy of its own end-of-tick view to that tick's gossip peers (the same Philox draw as its
// honest packets), whatever its transmit budgets say:
//   serf entry        leave intent (view Leaving/Left) or join intent (otherwise) at  status_time ∸ Δ
//   memberlist entry  its current state (alive/suspect/dead/left, same confirmer bucket) at  incarnation ∸ Δ
// (∸ saturates at 0).  The entries go through the ordinary inbox reduction, so a receiver whose view is even older
// accepts and re-gossips them like any other message.
// Anomaly flag of a SENDER u: set when some receiver v, up at the time the packet arrives, holds a view of that subject
// that is newer than the injected entry by at least Δ — serf: v knows the member and v.status_time ≥ sent_ltime + Δ;
// memberlist: v.incarnation ≥ sent_incarnation + Δ — judged on v's view at the end of the sending tick's node pass (what the
// packet meets on arrival, except for what a push-pull round of that very tick merges afterwards: the round runs after
// the verdicts, on the device as in the oracle).
#pragma once
#include "record.cuh"

namespace sfs {

struct ByzEntries { u32 serf_kind; u32 serf_lt; u32 ml_key; u32 ml_inc; bool any; };

__host__ __device__ inline u32 sat_sub(u32 a, u32 b) { return a > b ? a - b : 0u; }

__host__ __device__ inline ByzEntries byz_entries(const Rec& r, u32 delta) {
  ByzEntries e;
  e.any = (r.flags & 1u) != 0;                              // only subjects the node knows
  e.serf_kind = (r.status == ST_LEAVING || r.status == ST_LEFT) ? KIND_LEAVE : KIND_JOIN;
  e.serf_lt = sat_sub(r.st, delta);
  e.ml_inc = sat_sub(r.inc, delta);
  e.ml_key = (e.ml_inc << 6) | (r.mlstate << 4) | r.qfrom;
  return e;
}

// receiver-side judgement (evaluated by the sender's thread on the receiver's end-of-tick record)
__host__ __device__ inline bool byz_anomalous(const Rec& dst, const ByzEntries& e, u32 delta) {
  const bool serf = (dst.flags & 1u) && dst.st >= e.serf_lt + delta;
  const bool ml = dst.inc >= e.ml_inc + delta;
  return serf || ml;
}

}  // namespace sfs
