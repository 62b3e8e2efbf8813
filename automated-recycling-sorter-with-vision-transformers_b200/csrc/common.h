// Host-side plumbing shared by every translation unit of libvitk: error reporting that never
// throws across the C ABI, launch accounting, device properties, TMA descriptor encoding.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/vitk.h"

namespace vitk {

// Error codes are the VITK_* macros of include/vitk.h.

// printf-style; stores into a thread-local buffer, returns `code`.
int set_error(int code, const char* fmt, ...);
const char* last_error();

#define VITK_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::vitk::set_error(VITK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,      \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define VITK_CHECK_LAUNCH(name)                                                            \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::vitk::set_error(VITK_ERR_CUDA, "launch of %s failed: %s", name,     \
                               cudaGetErrorString(_e));                                    \
    ::vitk::count_launch();                                                                \
  } while (0)

#define VITK_REQUIRE(cond, ...)                                                            \
  do {                                                                                     \
    if (!(cond)) return ::vitk::set_error(VITK_ERR_INVALID, __VA_ARGS__);          \
  } while (0)

#define VITK_TRY(expr)                 \
  do {                                 \
    int _rc = (expr);                  \
    if (_rc != 0) return _rc;          \
  } while (0)

void count_launch();
long long launch_count();

// Optional per-kernel-class timing (bench.py's roofline leg): when enabled, every launcher brackets
// its kernel with CUDA events on the launching stream.  Off by default (zero overhead).
enum ProfKind : int { PROF_GEMM = 0, PROF_ATTN = 1, PROF_LN = 2, PROF_PATCH = 3, PROF_OTHER = 4,
                      PROF_OPT = 5, PROF_NKINDS = 6 };
bool profile_enabled();
void profile_enable(bool on);
// Sums (and clears) the recorded intervals: per kind milliseconds, work units (FLOPs for GEMM and
// attention, bytes for the HBM-bound kernels) and launch counts. Synchronises the recorded events.
int profile_collect(double* ms, double* work, long long* launches, int nkinds);

struct ProfileScope {
  ProfileScope(int kind, double work, cudaStream_t stream);
  ~ProfileScope();
  int idx_;
  cudaStream_t stream_;
};

// Sweep direction of the row-ordered kernels (LayerNorm rows, GEMM row tiles, attention items).
// Inside a SweepAlternation scope consecutive launches alternate between ascending and descending
// order, so every kernel starts on the rows its producer wrote LAST - the part of a 77-310 MB
// intermediate that is still resident in the 126 MB L2.  Outside a scope everything is ascending.
struct SweepAlternation {
  SweepAlternation();
  ~SweepAlternation();
};
int sweep_next();  // direction for the kernel being launched: 0 ascending, 1 descending

// Launch with programmatic stream serialization (see ptx.cuh: pdl_wait): the kernel's CTAs may be
// scheduled as soon as the preceding kernel's CTAs drain instead of after the whole grid has been
// retired and the next launch processed.  Off unless VITK_PDL=1 / vitk_set_pdl(1): measured on
// B200 (tests/ab_pdl.py, interleaved A/B) it changes the ViT-B/16 forward by < 1 % - the host runs
// far ahead of the device, so launch latency is already hidden.
bool pdl_enabled();
void set_pdl(int on);
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                       cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Cached per current device, minus the SMs set aside with reserve_sms (persistent kernels size
// their grids with it: a data-parallel backward leaves a few SMs to the NCCL kernels it overlaps).
int sm_count();
void reserve_sms(int n);
int device_cc();  // e.g. 100

// 2-D bf16 (or any 2-byte / 4-byte element) row-major tensor map: dims {inner, outer},
// row pitch in bytes, box {box_inner, box_outer}, 128-byte swizzle. Cached by value.
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner,
                 uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                 bool swizzle128 = true);

// 3-D variant: dims {inner, rows, batch}, pitches (bytes) of the rows and batch dimensions, box
// {box_inner, box_rows, 1}. Rows beyond `rows` are zero-filled on load and clipped on store, which
// keeps per-image token tiles from touching the neighbouring image.
int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t rows,
                 uint64_t batch, uint64_t row_pitch_bytes, uint64_t batch_pitch_bytes,
                 uint32_t box_inner, uint32_t box_rows);

}  // namespace vitk
