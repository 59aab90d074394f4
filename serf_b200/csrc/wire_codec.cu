// wire_codec.cu — serf's wire format behind the C ABI (layout and citations in wire.cuh).
//
// Single messages (Join, Leave, PushPull, with the message envelope) are encoded / decoded on the host — they are a few dozen
// bytes.  The bulk job is SerfDelegate::local_state (serf/delegate.rs:386-425): the push-pull message of a node lists every
// member it knows; producing it for all virtual nodes of a shard is an O(N·R) variable-length byte job, done on the device in
// three kernels: encoded lengths → exclusive scan → byte emission (one thread per node; a node's message is a few hundred bytes
// at most).  The inverse batch (decode n messages into per-node arrays) is the device half of merge_remote_state's parsing.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/serfsim.h"
#include "tick_kernel.cuh"
#include "wire.cuh"

using namespace sfs;
namespace w = sfs::wire;

struct serfsim;                                   // serfsim.cu
namespace sfs {
struct WireView { const uint4* rec; const u32* qword; const u64* node_state; const uint4* ue_state; const u32* subj; u32 n_local, stride, R; cudaStream_t stream; };
int serfsim_wire_view(const serfsim* h, WireView* out);          // serfsim.cu: the device arrays a batch encode reads
int serfsim_fail(int code, const char* msg);
}

namespace {

int werr(int rc) {
  switch (rc) {
    case w::E_TRUNCATED: return serfsim_fail(SERFSIM_E_INVAL, "wire: truncated message");
    case w::E_VARINT: return serfsim_fail(SERFSIM_E_INVAL, "wire: varint longer than 64 bits");
    case w::E_DUPLICATE: return serfsim_fail(SERFSIM_E_INVAL, "wire: duplicate field");
    case w::E_MISSING: return serfsim_fail(SERFSIM_E_INVAL, "wire: missing field");
    case w::E_WIRE_TYPE: return serfsim_fail(SERFSIM_E_INVAL, "wire: unknown wire type");
    case w::E_CAPACITY: return serfsim_fail(SERFSIM_E_INVAL, "wire: output capacity too small");
    case w::E_TYPE: return serfsim_fail(SERFSIM_E_INVAL, "wire: not a message of the requested type");
    default: return serfsim_fail(SERFSIM_E_INVAL, "wire: malformed message");
  }
}

// One node's local_state (serf/delegate.rs:386-425) from its views: the member table of a virtual node holds the tracked
// subjects it knows; left_members lists those it has as Left; the event clock comes from the user-event record (1 when user
// events are off — a fresh node, serf/base.rs:198-200), the query clock is not modelled (1).  No recent events are attached.
struct NodeState { u64 ltime, event_ltime; u32 known, left; };   // bit s of known / left: subject s
__device__ __forceinline__ u32 pp_payload_len(const WireView& v, u32 vl, NodeState* st_out, u64* sts) {
  NodeState st{};
  st.ltime = v.node_state[vl] & 0xffffffffull;
  st.event_ltime = v.ue_state ? v.ue_state[vl].x : 1u;
  u32 len = 1 + w::varint_len(st.ltime);
  for (u32 s = 0; s < v.R; ++s) {
    const size_t idx = (size_t)s * v.stride + vl;
    Rec r;
    unpack(v.rec[2 * idx], v.rec[2 * idx + 1], r);
    if (!(r.flags & FLAG_KNOWN)) continue;
    st.known |= 1u << s;
    if (sts) sts[s] = r.st;
    len += w::pp_status_entry_len(v.subj[s], r.st);
    if (r.status == ST_LEFT) { st.left |= 1u << s; }
  }
  for (u32 s = 0; s < v.R; ++s) if ((st.left >> s) & 1u) len += w::pp_left_entry_len(v.subj[s]);
  len += 1 + w::varint_len(st.event_ltime) + 1 + w::varint_len(1);
  *st_out = st;
  return len;
}
__global__ void pp_len_kernel(WireView v, u64* lens) {
  const u32 vl = blockIdx.x * blockDim.x + threadIdx.x;
  if (vl >= v.n_local) return;
  NodeState st;
  lens[vl] = w::envelope_len(pp_payload_len(v, vl, &st, nullptr));
}
// single-CTA exclusive scan over n u64 lengths (n ≤ a few 10 M: 1024 threads, a chunk each, then the chunk totals)
__global__ void __launch_bounds__(1024) pp_scan_kernel(const u64* lens, u64* offsets, u32 n) {
  __shared__ u64 part[1024];
  const u32 per = (n + 1023) / 1024, b = threadIdx.x * per, e = min(n, b + per);
  u64 s = 0;
  for (u32 i = b; i < e; ++i) s += lens[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) { u64 acc = 0; for (int i = 0; i < 1024; ++i) { const u64 t = part[i]; part[i] = acc; acc += t; } offsets[n] = acc; }
  __syncthreads();
  u64 acc = part[threadIdx.x];
  for (u32 i = b; i < e; ++i) { offsets[i] = acc; acc += lens[i]; }
}
__global__ void pp_emit_kernel(WireView v, const u64* offsets, u8* out, u64 cap) {
  const u32 vl = blockIdx.x * blockDim.x + threadIdx.x;
  if (vl >= v.n_local) return;
  NodeState st;
  u64 sts[MAX_SLOTS];
  const u32 pl = pp_payload_len(v, vl, &st, sts);
  if (offsets[vl] + w::envelope_len(pl) > cap) return;          // the host reports the shortfall from offsets[n]
  u8* p = out + offsets[vl];
  u32 o = 0;
  p[o++] = w::MSG_PUSH_PULL; o += w::varint_put(p + o, pl);                           // message.rs:397-428
  p[o++] = w::PP_LTIME; o += w::varint_put(p + o, st.ltime);                          // push_pull.rs:383-386
  for (u32 s = 0; s < v.R; ++s) if ((st.known >> s) & 1u) o += w::put_pp_status_entry(p + o, v.subj[s], sts[s]);   // :388-398
  for (u32 s = 0; s < v.R; ++s) if ((st.left >> s) & 1u) { p[o++] = w::PP_LEFT; o += w::varint_put(p + o, v.subj[s]); }   // :400-411
  p[o++] = w::PP_EVENT_LTIME; o += w::varint_put(p + o, st.event_ltime);              // :413-416 (no events attached, :418-428)
  p[o++] = w::PP_QUERY_LTIME; o += w::varint_put(p + o, 1);                           // :430-433
}

// PushPullMessageRef::decode (push_pull.rs:175-317) on one payload, arrays bounded by `cap` entries
__host__ __device__ int pp_decode_payload(const u8* p, size_t len, u64* ltime, u64* event_ltime, u64* query_ltime, u64* ids, u64* sts, u32 cap, u32* n_status,
                                          u64* left, u32 left_cap, u32* n_left, u32* n_events) {
  size_t o = 0;
  bool h_lt = false, h_ev = false, h_q = false;
  u32 ns = 0, nl = 0, ne = 0;
  while (o < len) {
    const u8 b = p[o];
    if (b == w::PP_LTIME || b == w::PP_EVENT_LTIME || b == w::PP_QUERY_LTIME) {
      bool& have = b == w::PP_LTIME ? h_lt : b == w::PP_EVENT_LTIME ? h_ev : h_q;
      if (have) return w::E_DUPLICATE;
      u64 v;
      const int r = w::varint_get(p + o + 1, len - o - 1, &v);
      if (r < 0) return r;
      *(b == w::PP_LTIME ? ltime : b == w::PP_EVENT_LTIME ? event_ltime : query_ltime) = v;
      have = true; o += 1 + r;
    } else if (b == w::PP_STATUS) {                          // one (id, status_time) tuple, length-delimited
      u64 tl;
      const int r = w::varint_get(p + o + 1, len - o - 1, &tl);
      if (r < 0) return r;
      if ((u64)(len - o - 1 - r) < tl) return w::E_TRUNCATED;
      const u8* t = p + o + 1 + r;
      size_t to = 0;
      u64 id = 0, st = 0; bool hk = false, hv = false;
      while (to < tl) {
        if (t[to] == w::TUPLE_KEY || t[to] == w::TUPLE_VALUE) {
          u64 v;
          const int r2 = w::varint_get(t + to + 1, (size_t)tl - to - 1, &v);
          if (r2 < 0) return r2;
          if (t[to] == w::TUPLE_KEY) { id = v; hk = true; } else { st = v; hv = true; }
          to += 1 + r2;
        } else {
          const long s = w::skip_field(t + to, (size_t)tl - to);
          if (s < 0) return (int)s;
          to += (size_t)s;
        }
      }
      if (!hk || !hv) return w::E_MISSING;
      if (ns >= cap) return w::E_CAPACITY;
      ids[ns] = id; sts[ns] = st; ++ns;
      o += 1 + r + (size_t)tl;
    } else if (b == w::PP_LEFT) {
      u64 v;
      const int r = w::varint_get(p + o + 1, len - o - 1, &v);
      if (r < 0) return r;
      if (nl >= left_cap) return w::E_CAPACITY;
      left[nl++] = v;
      o += 1 + r;
    } else {
      if (b == w::PP_EVENTS) ++ne;                           // recent events: counted and skipped (not part of the membership path)
      const long s = w::skip_field(p + o, len - o);
      if (s < 0) return (int)s;
      o += (size_t)s;
    }
  }
  if (!h_lt || !h_ev || !h_q) return w::E_MISSING;            // push_pull.rs:292-316
  *n_status = ns; *n_left = nl; if (n_events) *n_events = ne;
  return w::OK;
}
__global__ void pp_decode_kernel(const u8* buf, const u64* offsets, u32 n, u32 cap, u64* ltime, u64* ids, u64* sts, u32* n_status, u32* left_mask_err) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8* m = buf + offsets[i];
  const size_t len = (size_t)(offsets[i + 1] - offsets[i]);
  u8 type = 0; size_t po = 0, pl = 0;
  int rc = w::open_envelope(m, len, &type, &po, &pl);
  if (rc == w::OK && type != w::MSG_PUSH_PULL) rc = w::E_TYPE;
  u64 ev, q, left[MAX_SLOTS];
  u32 ns = 0, nl = 0;
  if (rc == w::OK) rc = pp_decode_payload(m + po, pl, ltime + i, &ev, &q, ids + (size_t)i * cap, sts + (size_t)i * cap, cap, &ns, left, MAX_SLOTS, &nl, nullptr);
  n_status[i] = rc == w::OK ? ns : 0;
  left_mask_err[i] = rc == w::OK ? nl : 0x80000000u | (u32)(-rc);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

size_t serfsim_wire_encoded_len_intent(const serfsim_wire_intent_t* m) {
  if (!m) return 0;
  return w::envelope_len(m->type == SERFSIM_WIRE_JOIN ? w::join_payload_len(m->ltime, m->id) : w::leave_payload_len(m->ltime, m->id, m->prune != 0));
}

int serfsim_wire_encode_intent(const serfsim_wire_intent_t* m, uint8_t* buf, size_t cap, size_t* len) {
  if (!m || !buf || !len) return serfsim_fail(SERFSIM_E_INVAL, "null argument");
  if (m->type != SERFSIM_WIRE_JOIN && m->type != SERFSIM_WIRE_LEAVE) return serfsim_fail(SERFSIM_E_INVAL, "wire: intent type must be SERFSIM_WIRE_JOIN or SERFSIM_WIRE_LEAVE");
  const size_t need = serfsim_wire_encoded_len_intent(m);
  *len = need;
  if (cap < need) return werr(w::E_CAPACITY);                 // EncodeError::insufficient_buffer: the needed size is reported in *len
  const u32 n = m->type == SERFSIM_WIRE_JOIN ? w::put_join(buf, m->ltime, m->id) : w::put_leave(buf, m->ltime, m->id, m->prune != 0);
  return n == need ? 0 : serfsim_fail(SERFSIM_E_INVAL, "wire: internal length mismatch");
}

int serfsim_wire_message_type(const uint8_t* buf, size_t len, uint32_t* type) {
  if (!buf || !type) return serfsim_fail(SERFSIM_E_INVAL, "null argument");
  u8 t = 0; size_t po = 0, pl = 0;
  const int rc = w::open_envelope(buf, len, &t, &po, &pl);
  if (rc) return werr(rc);
  *type = t == w::MSG_LEAVE ? SERFSIM_WIRE_LEAVE : t == w::MSG_JOIN ? SERFSIM_WIRE_JOIN : SERFSIM_WIRE_PUSH_PULL;
  return 0;
}

int serfsim_wire_decode_intent(const uint8_t* buf, size_t len, serfsim_wire_intent_t* out) {
  if (!buf || !out) return serfsim_fail(SERFSIM_E_INVAL, "null argument");
  u8 t = 0; size_t po = 0, pl = 0;
  int rc = w::open_envelope(buf, len, &t, &po, &pl);
  if (rc) return werr(rc);
  if (t != w::MSG_JOIN && t != w::MSG_LEAVE) return werr(w::E_TYPE);
  w::Intent in{};
  rc = w::get_intent(buf + po, pl, t == w::MSG_LEAVE, &in);
  if (rc) return werr(rc);
  out->type = t == w::MSG_LEAVE ? SERFSIM_WIRE_LEAVE : SERFSIM_WIRE_JOIN; out->prune = in.prune ? 1u : 0u; out->ltime = in.ltime; out->id = in.id;
  return 0;
}

int serfsim_wire_encode_push_pull(const serfsim_wire_push_pull_t* m, uint8_t* buf, size_t cap, size_t* len) {
  if (!m || !len || (m->n_status && (!m->status_ids || !m->status_ltimes)) || (m->n_left && !m->left_ids)) return serfsim_fail(SERFSIM_E_INVAL, "null argument");
  size_t pl = 1 + w::varint_len(m->ltime);
  for (u32 i = 0; i < m->n_status; ++i) pl += w::pp_status_entry_len(m->status_ids[i], m->status_ltimes[i]);
  for (u32 i = 0; i < m->n_left; ++i) pl += w::pp_left_entry_len(m->left_ids[i]);
  pl += 1 + w::varint_len(m->event_ltime) + 1 + w::varint_len(m->query_ltime);
  if (pl > 0xffffffffull) return serfsim_fail(SERFSIM_E_INVAL, "wire: message too large");      // EncodeError::TooLarge, message.rs:410-412
  const size_t need = 1 + w::varint_len(pl) + pl;
  *len = need;
  if (!buf || cap < need) return werr(w::E_CAPACITY);
  size_t o = 0;
  buf[o++] = w::MSG_PUSH_PULL; o += w::varint_put(buf + o, pl);
  buf[o++] = w::PP_LTIME; o += w::varint_put(buf + o, m->ltime);
  for (u32 i = 0; i < m->n_status; ++i) o += w::put_pp_status_entry(buf + o, m->status_ids[i], m->status_ltimes[i]);
  for (u32 i = 0; i < m->n_left; ++i) { buf[o++] = w::PP_LEFT; o += w::varint_put(buf + o, m->left_ids[i]); }
  buf[o++] = w::PP_EVENT_LTIME; o += w::varint_put(buf + o, m->event_ltime);
  buf[o++] = w::PP_QUERY_LTIME; o += w::varint_put(buf + o, m->query_ltime);
  return o == need ? 0 : serfsim_fail(SERFSIM_E_INVAL, "wire: internal length mismatch");
}

int serfsim_wire_decode_push_pull(const uint8_t* buf, size_t len, serfsim_wire_push_pull_t* out) {
  if (!buf || !out) return serfsim_fail(SERFSIM_E_INVAL, "null argument");
  u8 t = 0; size_t po = 0, pl = 0;
  int rc = w::open_envelope(buf, len, &t, &po, &pl);
  if (rc) return werr(rc);
  if (t != w::MSG_PUSH_PULL) return werr(w::E_TYPE);
  u32 ns = 0, nl = 0, ne = 0;
  rc = pp_decode_payload(buf + po, pl, &out->ltime, &out->event_ltime, &out->query_ltime, out->status_ids, out->status_ltimes, out->n_status, &ns,
                         out->left_ids, out->n_left, &nl, &ne);
  if (rc) return werr(rc);
  out->n_status = ns; out->n_left = nl; out->n_events_skipped = ne;
  return 0;
}

// SerfDelegate::local_state of every node of the shard, on the device.  offsets: [count + 1] byte offsets into `out`
// (offsets[count] = total).  With out == NULL or cap too small only the offsets are produced and SERFSIM_E_INVAL is returned
// with *total set, so that the caller can size the buffer.
int serfsim_wire_local_state_batch(serfsim_t* h, uint8_t* out, size_t cap, uint64_t* offsets, size_t* total) {
  if (!h || !offsets || !total) return serfsim_fail(SERFSIM_E_INVAL, "null argument");
  WireView v{};
  int rc = serfsim_wire_view(h, &v);
  if (rc) return rc;
  const u32 n = v.n_local;
  u64 *d_len = nullptr, *d_off = nullptr; u8* d_out = nullptr;
  auto cleanup = [&]() { cudaFree(d_len); cudaFree(d_off); cudaFree(d_out); };
#define CW(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return serfsim_fail(SERFSIM_E_CUDA, cudaGetErrorString(e_)); } } while (0)
  CW(cudaMalloc(&d_len, (size_t)n * 8)); CW(cudaMalloc(&d_off, ((size_t)n + 1) * 8));
  SFS_LAUNCH((n + 255) / 256, 256, 0, v.stream, pp_len_kernel)(v, d_len);
  SFS_LAUNCH(1, 1024, 0, v.stream, pp_scan_kernel)(d_len, d_off, n);
  CW(cudaMemcpyAsync(offsets, d_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, v.stream));
  CW(cudaStreamSynchronize(v.stream));
  *total = (size_t)offsets[n];
  if (!out || cap < *total) { cleanup(); return werr(w::E_CAPACITY); }
  CW(cudaMalloc(&d_out, *total ? *total : 1));
  SFS_LAUNCH((n + 255) / 256, 256, 0, v.stream, pp_emit_kernel)(v, d_off, d_out, (u64)*total);
  CW(cudaMemcpyAsync(out, d_out, *total, cudaMemcpyDeviceToHost, v.stream));
  CW(cudaStreamSynchronize(v.stream));
  CW(cudaGetLastError());
  cleanup();
  return 0;
}

// The inverse batch on the device: n push-pull messages (concatenated, offsets[n + 1]) → per message the Lamport clock, up to
// `cap` (id, status_time) entries and their count.  A malformed message fails the call (index in the error text).
int serfsim_wire_decode_batch(serfsim_t* h, const uint8_t* buf, const uint64_t* offsets, uint32_t n, uint32_t cap, uint64_t* ltime, uint64_t* ids, uint64_t* status_ltimes, uint32_t* n_status) {
  if (!h || !buf || !offsets || !ltime || !ids || !status_ltimes || !n_status || !cap) return serfsim_fail(SERFSIM_E_INVAL, "null argument");
  WireView v{};
  int rc = serfsim_wire_view(h, &v);
  if (rc) return rc;
  const size_t total = (size_t)offsets[n];
  u8* d_buf = nullptr; u64 *d_off = nullptr, *d_lt = nullptr, *d_ids = nullptr, *d_sts = nullptr; u32 *d_ns = nullptr, *d_err = nullptr;
  auto cleanup = [&]() { cudaFree(d_buf); cudaFree(d_off); cudaFree(d_lt); cudaFree(d_ids); cudaFree(d_sts); cudaFree(d_ns); cudaFree(d_err); };
  CW(cudaMalloc(&d_buf, total ? total : 1)); CW(cudaMalloc(&d_off, ((size_t)n + 1) * 8)); CW(cudaMalloc(&d_lt, (size_t)n * 8 + 8));
  CW(cudaMalloc(&d_ids, (size_t)n * cap * 8 + 8)); CW(cudaMalloc(&d_sts, (size_t)n * cap * 8 + 8)); CW(cudaMalloc(&d_ns, (size_t)n * 4 + 4)); CW(cudaMalloc(&d_err, (size_t)n * 4 + 4));
  CW(cudaMemcpyAsync(d_buf, buf, total, cudaMemcpyHostToDevice, v.stream));
  CW(cudaMemcpyAsync(d_off, offsets, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, v.stream));
  if (n) SFS_LAUNCH((n + 127) / 128, 128, 0, v.stream, pp_decode_kernel)(d_buf, d_off, n, cap, d_lt, d_ids, d_sts, d_ns, d_err);
  std::vector<u32> err(n);
  CW(cudaMemcpyAsync(ltime, d_lt, (size_t)n * 8, cudaMemcpyDeviceToHost, v.stream));
  CW(cudaMemcpyAsync(ids, d_ids, (size_t)n * cap * 8, cudaMemcpyDeviceToHost, v.stream));
  CW(cudaMemcpyAsync(status_ltimes, d_sts, (size_t)n * cap * 8, cudaMemcpyDeviceToHost, v.stream));
  CW(cudaMemcpyAsync(n_status, d_ns, (size_t)n * 4, cudaMemcpyDeviceToHost, v.stream));
  CW(cudaMemcpyAsync(err.data(), d_err, (size_t)n * 4, cudaMemcpyDeviceToHost, v.stream));
  CW(cudaStreamSynchronize(v.stream));
  CW(cudaGetLastError());
  cleanup();
#undef CW
  for (u32 i = 0; i < n; ++i)
    if (err[i] & 0x80000000u) { werr(-(int)(err[i] & 0xffffu)); return serfsim_fail(SERFSIM_E_INVAL, (std::string("wire: message ") + std::to_string(i) + ": " + serfsim_last_error()).c_str()); }
  return 0;
}

}  // extern "C"
#pragma GCC visibility pop
