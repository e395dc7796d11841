// api.cu — host plumbing of the C-ABI: error strings, device gate, TMA tensor-map encoding.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace nans {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device: this library has no CPU path");
    return NANS_ERR_DEVICE;
  }
  static thread_local int cached_dev = -1, cached_major = 0;
  if (cached_dev != dev) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaDeviceGetAttribute failed");
      return NANS_ERR_DEVICE;
    }
    cached_dev = dev;
    cached_major = major;
  }
  if (cached_major != 10) {
    set_error("device %d is compute capability %d.x; this library is sm_100a only", dev,
              cached_major);
    return NANS_ERR_DEVICE;
  }
  return NANS_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) {
      cached = n;
      cached_dev = dev;
    }
  }
  return cached;
}

static PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;  // benign race: same value from every thread
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    } else {
      cudaGetLastError();
    }
  }
  return fn;
}

int make_tmap_16b(CUtensorMap* out, const void* base, int feat_dtype, uint64_t rows,
                  uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return NANS_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld_elems * 2) % 16 != 0) {
    set_error("feature matrix must be 16-byte aligned with a row pitch multiple of 16 bytes");
    return NANS_ERR_ARG;
  }
  if (box_rows == 0 || box_rows > 256) {
    set_error("internal: TMA box rows %u out of range", box_rows);
    return NANS_ERR_ARG;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt =
      feat_dtype == NANS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu)",
              static_cast<int>(r), (unsigned long long)rows, (unsigned long long)cols,
              (unsigned long long)ld_elems);
    return NANS_ERR_CUDA;
  }
  return NANS_OK;
}

}  // namespace nans

extern "C" {

int nans_version(void) { return NANS_VERSION; }

const char* nans_last_error(void) { return nans::g_err; }

int nans_device_check(void) { return nans::check_device(); }

}  // extern "C"
