#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace refid {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("REFID_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
int set_pdl(int enable) {
  const int prev = pdl_enabled() ? 1 : 0;
  g_pdl = enable ? 1 : 0;
  return prev;
}


// ---- abort flags: one device copy per translation unit + one pinned host word (see common.cuh) ----
namespace {
struct AbortReg {
  const void* flag;
  const void* host_ptr;
};
AbortReg g_abort_regs[32];
int g_abort_nregs = 0;
volatile unsigned int* g_abort_host_word = nullptr;  // pinned + mapped
bool g_abort_published = false;
}  // namespace

void register_abort_flag(const void* flag_symbol, const void* host_ptr_symbol) {
  if (g_abort_nregs < 32) g_abort_regs[g_abort_nregs++] = AbortReg{flag_symbol, host_ptr_symbol};
}

// Hands every translation unit the device address of the pinned host word (once per process, first GPU call).
static int publish_abort_word() {
  if (g_abort_published) return 0;
  unsigned int* h = nullptr;
  REFID_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable));
  *h = 0;
  unsigned int* d = nullptr;
  REFID_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0));
  for (int i = 0; i < g_abort_nregs; ++i)
    REFID_CUDA_CHECK(cudaMemcpyToSymbol(g_abort_regs[i].host_ptr, &d, sizeof(d)));
  g_abort_host_word = h;
  g_abort_published = true;
  return 0;
}

unsigned int abort_pending() {
  if (!g_abort_published && publish_abort_word()) return 0xFFFFFFFFu;
  return *g_abort_host_word;
}

int read_and_clear_abort_flag(cudaStream_t stream, unsigned int* out) {
  unsigned int first = 0, z = 0;
  if (publish_abort_word()) return 1;
  REFID_CUDA_CHECK(cudaStreamSynchronize(stream));
  REFID_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < g_abort_nregs; ++i) {
    unsigned int v = 0;
    REFID_CUDA_CHECK(cudaMemcpyFromSymbol(&v, g_abort_regs[i].flag, sizeof(v)));
    if (v) {
      REFID_CUDA_CHECK(cudaMemcpyToSymbol(g_abort_regs[i].flag, &z, sizeof(z)));
      if (!first) first = v;
    }
  }
  if (!first) first = *g_abort_host_word;
  *g_abort_host_word = 0;
  *out = first;
  return 0;
}

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time libcuda dependency.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  if (inner_bytes == 128) return CU_TENSOR_MAP_SWIZZLE_128B;
  if (inner_bytes == 64) return CU_TENSOR_MAP_SWIZZLE_64B;
  if (inner_bytes == 32) return CU_TENSOR_MAP_SWIZZLE_32B;
  return CU_TENSOR_MAP_SWIZZLE_NONE;
}

int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int pitch, int C, int y0, int x0, int step,
                 int boxC, int TW, int TH, int TN) {
  EncodeTiledFn enc = get_encode();
  REFID_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  REFID_REQUIRE(boxC == 64 || boxC == 32, "make_act_map: boxC must be 32 or 64 (got %d)", boxC);
  REFID_REQUIRE(C % boxC == 0, "make_act_map: C=%d not a multiple of boxC=%d", C, boxC);
  const int Ws = (W - x0 + step - 1) / step, Hs = (H - y0 + step - 1) / step;
  const char* b = static_cast<const char*>(base) + ((size_t)y0 * W + x0) * pitch * 2;
  REFID_REQUIRE((reinterpret_cast<uintptr_t>(b) & 15) == 0, "make_act_map: base not 16B aligned");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 2 * step, (cuuint64_t)W * pitch * 2 * step, (cuuint64_t)H * W * pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(b), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(boxC * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REFID_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(act) failed: %d (N%d H%d W%d C%d pitch%d step%d box %d,%d,%d,%d)",
                (int)r, N, H, W, C, pitch, step, boxC, TW, TH, TN);
  return 0;
}

int make_mat_map(CUtensorMap* m, const void* base, long rows, long cols, int boxCols, int boxRows) {
  EncodeTiledFn enc = get_encode();
  REFID_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  REFID_REQUIRE(boxCols == 64 || boxCols == 32, "make_mat_map: boxCols must be 32 or 64");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxCols, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(boxCols * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REFID_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(mat) failed: %d (rows %ld cols %ld box %d,%d)", (int)r, rows,
                cols, boxCols, boxRows);
  return 0;
}

}  // namespace refid
