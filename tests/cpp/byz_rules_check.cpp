// Host-compiled view of the byzantine-injector rules the CUDA kernel runs (serf_b200/csrc/byz.cuh), so that
// tests/test_byzantine.py can compare them with the oracle's definition on random records without a GPU.
// Test infrastructure only.
#include <cstdint>
#include <cstring>
struct uint4 { unsigned x, y, z, w; };
#define __host__
#define __device__
#include "../../serf_b200/csrc/byz.cuh"

using namespace sfs;

static Rec load(const void* rec32) {
  uint4 a, b;
  memcpy(&a, rec32, 16);
  memcpy(&b, (const char*)rec32 + 16, 16);
  Rec r;
  unpack(a, b, r);
  return r;
}

extern "C" __attribute__((visibility("default")))
void byzcheck_entries(const void* rec32, u32 delta, u32* out /*any, serf_kind, serf_lt, ml_key, ml_inc*/) {
  const ByzEntries e = byz_entries(load(rec32), delta);
  out[0] = e.any; out[1] = e.serf_kind; out[2] = e.serf_lt; out[3] = e.ml_key; out[4] = e.ml_inc;
}

extern "C" __attribute__((visibility("default")))
int byzcheck_anomalous(const void* dst32, u32 serf_lt, u32 ml_inc, u32 delta) {
  ByzEntries e{};
  e.serf_lt = serf_lt; e.ml_inc = ml_inc;
  return byz_anomalous(load(dst32), e, delta);
}
