// Host-side plumbing of libb200diff.so: error text, launch counter, TMA tensor-map encoding through the
// driver entry point (resolved at run time so the library links against cudart only).
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/b200diff.h"

namespace b200 {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B200_PDL");
    on = (e && atoi(e) == 1) ? 1 : 0;   // measured on B200 (DDIM-50 graph replay): 958 vs 969 images/s, no gain -> off by default
  }
  return on != 0;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return 1001;
    }
    g_encode = reinterpret_cast<encode_tiled_fn>(fn);
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
  CUresult r = g_encode(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return 1002;
  }
  return 0;
}

}  // namespace b200

extern "C" int b200_version(void) { return B200DIFF_VERSION; }
extern "C" const char* b200_last_error(void) { return b200::g_err; }
extern "C" long long b200_launch_count(void) { return b200::g_launch_count; }
