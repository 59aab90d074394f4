// wire_oracle.cpp — CPU restatement of serf's wire format for Join / Leave / PushPull messages.
// TEST INFRASTRUCTURE: the checker of serf_b200/csrc/wire_codec.cu; nothing under serf_b200/ links, loads or calls it.
//
// Follows serf-core/src/types: message.rs:17-47 (message tags), :397-428 (encode_message), :484-491 (encoded_message_len),
// :507-692 (decode_message); join.rs:8-10, :54-158; leave.rs:8-13, :56-195; push_pull.rs:100-114, :175-317, :319-450;
// clock.rs:96-119 (LamportTime is a u64 on the wire).  Ids are u64 (the `..U64` instantiations of types/tests.rs:49-62).
//
// PARITY UNPINNED AT BYTE LEVEL: the primitives those files import — merge / skip / WireType / varint / TupleEncoder /
// encode_length_delimited — live in the external crate memberlist-core 0.8.1 (memberlist_core::proto; Cargo.toml:39-41, no
// Cargo.lock, not vendored) and the reference tree holds no golden bytes, only the round-trip property of types/tests.rs:8-25.
// They are restated from the protobuf conventions that crate follows: tag byte = tag << 3 | wire type (message tags reach 10,
// so the tag cannot sit in the low three bits); wire types Byte 0, Varint 1, LengthDelimited 2, Fixed32 3, Fixed64 4; LEB128
// varints; a bool is one byte; a Varint-typed id gets no length prefix from encode_length_delimited; TupleEncoder(k, v) is a
// map entry [tag 1: k][tag 2: v].  What IS pinned by the tree: the field sets, tags, order of emission, which fields are
// optional (prune is written only when true), duplicate- and missing-field errors, skipping of unknown fields, and
// decode(encode(m)) == m with consumed == encoded_len for every message.
#include <cstdint>
#include <cstring>
#include <vector>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
#define ORC extern "C" __attribute__((visibility("default")))

namespace {

enum { BYTE = 0, VARINT = 1, LEN = 2, FIXED32 = 3, FIXED64 = 4 };
inline u8 merge(int wire, int tag) { return (u8)((tag << 3) | wire); }

void put_varint(std::vector<u8>& b, u64 v) { while (v >= 0x80) { b.push_back((u8)(v | 0x80)); v >>= 7; } b.push_back((u8)v); }
bool get_varint(const u8* p, size_t n, size_t& o, u64& v) {
  v = 0;
  for (int i = 0; i < 10; ++i) {
    if (o >= n) return false;
    const u8 c = p[o++];
    if (i == 9 && c > 1) return false;
    v |= (u64)(c & 0x7f) << (7 * i);
    if (!(c & 0x80)) return true;
  }
  return false;
}
bool skip(const u8* p, size_t n, size_t& o) {                // memberlist_core::proto::utils::skip, restated
  const int wire = p[o] & 7;
  ++o;
  u64 v;
  switch (wire) {
    case BYTE: if (o + 1 > n) return false; o += 1; return true;
    case VARINT: return get_varint(p, n, o, v);
    case LEN: if (!get_varint(p, n, o, v) || n - o < v) return false; o += (size_t)v; return true;
    case FIXED32: if (n - o < 4) return false; o += 4; return true;
    case FIXED64: if (n - o < 8) return false; o += 8; return true;
    default: return false;
  }
}
std::vector<u8> envelope(u8 id, const std::vector<u8>& payload) {         // encode_message, message.rs:397-428
  std::vector<u8> b;
  b.push_back(id);
  put_varint(b, (u32)payload.size());
  b.insert(b.end(), payload.begin(), payload.end());
  return b;
}
// returns 0 ok, -3 duplicate, -4 missing, -1 malformed
int open(const u8* p, size_t n, u8& type, size_t& po, size_t& pl) {        // decode_message, message.rs:507-692
  size_t o = 0;
  bool have = false;
  while (o < n) {
    const u8 b = p[o];
    if (b == merge(LEN, 1) || b == merge(LEN, 2) || b == merge(LEN, 3)) {
      if (have) return -3;
      ++o;
      u64 len;
      if (!get_varint(p, n, o, len) || n - o < len) return -1;
      type = b; po = o; pl = (size_t)len; have = true;
      o += (size_t)len;
    } else if (!skip(p, n, o)) return -1;
  }
  return have ? 0 : -4;
}
int copy_out(const std::vector<u8>& b, u8* out, size_t cap, size_t* len) { *len = b.size(); if (cap < b.size()) return -6; memcpy(out, b.data(), b.size()); return 0; }

}  // namespace

// type 1 = Leave (leave.rs:141-195), 2 = Join (join.rs:131-158)
ORC int oracle_wire_encode_intent(u32 type, u64 ltime, u64 id, int prune, u8* out, size_t cap, size_t* len) {
  std::vector<u8> p;
  if (type == 2) {
    p.push_back(merge(VARINT, 1)); put_varint(p, ltime);
    p.push_back(merge(VARINT, 2)); put_varint(p, id);
    return copy_out(envelope(merge(LEN, 2), p), out, cap, len);
  }
  p.push_back(merge(VARINT, 1)); put_varint(p, ltime);
  if (prune) { p.push_back(merge(BYTE, 2)); p.push_back(1); }
  p.push_back(merge(VARINT, 3)); put_varint(p, id);
  return copy_out(envelope(merge(LEN, 1), p), out, cap, len);
}
ORC int oracle_wire_decode_intent(const u8* buf, size_t n, u32* type, u64* ltime, u64* id, int* prune) {
  u8 t; size_t po, pl;
  int rc = open(buf, n, t, po, pl);
  if (rc) return rc;
  if (t == merge(LEN, 3)) return -7;
  const bool leave = t == merge(LEN, 1);
  const u8* p = buf + po;
  size_t o = 0;
  bool hl = false, hi = false, hp = false;
  *prune = 0;
  while (o < pl) {
    const u8 b = p[o];
    if (b == merge(VARINT, 1)) { if (hl) return -3; ++o; if (!get_varint(p, pl, o, *ltime)) return -1; hl = true; }
    else if (leave && b == merge(BYTE, 2)) { if (hp) return -3; if (o + 2 > pl) return -1; *prune = p[o + 1] != 0; o += 2; hp = true; }
    else if (b == merge(VARINT, leave ? 3 : 2)) { if (!leave && hi) return -3; ++o; if (!get_varint(p, pl, o, *id)) return -1; hi = true; }
    else if (!skip(p, pl, o)) return -1;
  }
  if (!hl || !hi) return -4;
  *type = leave ? 1 : 2;
  return 0;
}
// push_pull.rs:361-450
ORC int oracle_wire_encode_push_pull(u64 ltime, const u64* ids, const u64* sts, u32 n_status, const u64* left, u32 n_left, u64 event_ltime, u64 query_ltime,
                                     u8* out, size_t cap, size_t* len) {
  std::vector<u8> p;
  p.push_back(merge(VARINT, 1)); put_varint(p, ltime);
  for (u32 i = 0; i < n_status; ++i) {
    std::vector<u8> t;
    t.push_back(merge(VARINT, 1)); put_varint(t, ids[i]);
    t.push_back(merge(VARINT, 2)); put_varint(t, sts[i]);
    p.push_back(merge(LEN, 2)); put_varint(p, t.size()); p.insert(p.end(), t.begin(), t.end());
  }
  for (u32 i = 0; i < n_left; ++i) { p.push_back(merge(VARINT, 3)); put_varint(p, left[i]); }
  p.push_back(merge(VARINT, 4)); put_varint(p, event_ltime);
  p.push_back(merge(VARINT, 6)); put_varint(p, query_ltime);
  return copy_out(envelope(merge(LEN, 3), p), out, cap, len);
}
// push_pull.rs:175-317; capacities in *n_status / *n_left, entries out
ORC int oracle_wire_decode_push_pull(const u8* buf, size_t n, u64* ltime, u64* ids, u64* sts, u32* n_status, u64* left, u32* n_left, u64* event_ltime, u64* query_ltime, u32* n_events) {
  u8 t; size_t po, pl;
  int rc = open(buf, n, t, po, pl);
  if (rc) return rc;
  if (t != merge(LEN, 3)) return -7;
  const u8* p = buf + po;
  size_t o = 0;
  bool hl = false, he = false, hq = false;
  u32 ns = 0, nl = 0, ne = 0;
  while (o < pl) {
    const u8 b = p[o];
    if (b == merge(VARINT, 1)) { if (hl) return -3; ++o; if (!get_varint(p, pl, o, *ltime)) return -1; hl = true; }
    else if (b == merge(VARINT, 4)) { if (he) return -3; ++o; if (!get_varint(p, pl, o, *event_ltime)) return -1; he = true; }
    else if (b == merge(VARINT, 6)) { if (hq) return -3; ++o; if (!get_varint(p, pl, o, *query_ltime)) return -1; hq = true; }
    else if (b == merge(LEN, 2)) {
      ++o;
      u64 tl;
      if (!get_varint(p, pl, o, tl) || pl - o < tl) return -1;
      const u8* q = p + o;
      size_t qo = 0;
      u64 id = 0, st = 0; bool hk = false, hv = false;
      while (qo < tl) {
        if (q[qo] == merge(VARINT, 1)) { ++qo; if (!get_varint(q, (size_t)tl, qo, id)) return -1; hk = true; }
        else if (q[qo] == merge(VARINT, 2)) { ++qo; if (!get_varint(q, (size_t)tl, qo, st)) return -1; hv = true; }
        else if (!skip(q, (size_t)tl, qo)) return -1;
      }
      if (!hk || !hv) return -4;
      if (ns >= *n_status) return -6;
      ids[ns] = id; sts[ns] = st; ++ns;
      o += (size_t)tl;
    }
    else if (b == merge(VARINT, 3)) { ++o; u64 v; if (!get_varint(p, pl, o, v)) return -1; if (nl >= *n_left) return -6; left[nl++] = v; }
    else { if (b == merge(LEN, 5)) ++ne; if (!skip(p, pl, o)) return -1; }
  }
  if (!hl || !he || !hq) return -4;
  *n_status = ns; *n_left = nl; *n_events = ne;
  return 0;
}
