// The module's parameter tensors <-> the engine's flat fp32 "gradient layout" vector, as ONE gather (forward) and ONE
// scatter (backward) over a table of entries, instead of ~300 small framework ops per training step (round 1: building the
// flat vector and back-propagating through it were ~140 + ~400 ATen launches and most of the 7.7 ms of host time per step
// that remained once the network itself replayed from CUDA graphs).
//   mode 0: contiguous copy of n = taps*R*Cc floats (biases, raw vectors, already-folded sites)
//   mode 1: conv weight (Cc, R, taps) [PyTorch (Cout,Cin,kh,kw); ConvTranspose (Cin,Cout,2,2) with R = Cout, Cc = Cin]
//           -> [tap][R][Cc]:  flat[off + (t*R + r)*Cc + c] = src[(c*R + r)*taps + t]
// Pure data movement: 4 B read + 4 B written per element.
#include "../../include/refid_b200.h"
#include "common.cuh"

namespace refid {
namespace {

constexpr int kFlatBatch = 96;  // entries per launch (the table travels as a kernel argument)
struct FlatTable {
  refid_flat_entry e[kFlatBatch];
};

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_flat(const __grid_constant__ FlatTable tab, float* __restrict__ flat) {
  const refid_flat_entry& e = tab.e[blockIdx.y];
  const long n = (long)e.taps * e.R * e.Cc;
  float* p = static_cast<float*>(const_cast<void*>(e.ptr));
  float* f = flat + e.flat_off;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long j = i;
    if (e.mode == 1) {
      const int c = (int)(i % e.Cc);
      const long tr = i / e.Cc;
      const int r = (int)(tr % e.R), t = (int)(tr / e.R);
      j = ((long)c * e.R + r) * e.taps + t;
    }
    if (SCATTER) p[j] = f[i];
    else f[i] = p[j];
  }
}

template <bool SCATTER>
int run_flat(const refid_flat_entry* entries, int n, float* flat, cudaStream_t st) {
  for (int i0 = 0; i0 < n; i0 += kFlatBatch) {
    FlatTable tab;
    memset(&tab, 0, sizeof(tab));
    const int m = n - i0 < kFlatBatch ? n - i0 : kFlatBatch;
    long mx = 1;
    for (int i = 0; i < m; ++i) {
      tab.e[i] = entries[i0 + i];
      REFID_REQUIRE(tab.e[i].ptr && tab.e[i].flat_off >= 0 && tab.e[i].taps > 0 && tab.e[i].R > 0 && tab.e[i].Cc > 0 &&
                        (tab.e[i].mode == 0 || tab.e[i].mode == 1),
                    "flat table entry %d is malformed", i0 + i);
      const long ne = (long)tab.e[i].taps * tab.e[i].R * tab.e[i].Cc;
      if (ne > mx) mx = ne;
    }
    long bx = (mx + 256 * 8 - 1) / (256 * 8);
    if (bx > 64) bx = 64;
    k_flat<SCATTER><<<dim3((unsigned)bx, (unsigned)m), 256, 0, st>>>(tab, flat);
    REFID_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

}  // namespace
}  // namespace refid

extern "C" {

int refid_flat_gather(const refid_flat_entry* entries, int n, float* flat, long flat_floats, void* stream) {
  using namespace refid;
  REFID_REQUIRE(entries && flat && n > 0 && flat_floats > 0, "refid_flat_gather: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  REFID_CUDA_CHECK(cudaMemsetAsync(flat, 0, (size_t)flat_floats * 4, st));  // alignment gaps and padded rows stay zero
  return run_flat<false>(entries, n, flat, st);
}

int refid_flat_scatter(const refid_flat_entry* entries, int n, const float* gflat, void* stream) {
  using namespace refid;
  REFID_REQUIRE(entries && gflat && n > 0, "refid_flat_scatter: bad argument");
  return run_flat<true>(entries, n, const_cast<float*>(gflat), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
