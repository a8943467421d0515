// Host-side launchers for the conv stack and the tensor-core GEMM, plus the CUDA-core first layer.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "sed_conv.cuh"
#include "sed_kernels.h"

namespace sed {

// ---------------------------------------------------------------------------------------------
// error string (thread-local)
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency,
// so the library loads on hosts without a driver for symbol checks)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_map(CUtensorMap* m, int dtype, int rank, void* base, const uint64_t* dims, const uint64_t* strides_b,
                    const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SED_ERR_DRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_b[i - 1];
  }
  CUresult r = enc(m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, gdim,
                   gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
    return SED_ERR_DRIVER;
  }
  return SED_OK;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// conv_block1.conv1 (Cin = 1): CUDA cores, float32 in, 16-bit NHWC out.  models.py:128 (first conv)
// Each thread owns 8 output channels (weights in registers) and walks over pixels; a warp writes
// 4 pixels x 128 B contiguous per store instruction.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv_first_kernel(const float* __restrict__ x, int H, int W, const float* __restrict__ w9,
                  const float* __restrict__ scale, const float* __restrict__ shift, T* __restrict__ out) {
  constexpr int ROWS = 8;
  __shared__ float s_in[ROWS + 2][72];  // W == 64 plus one zero column each side
  const int n = blockIdx.y;
  const int h_base = blockIdx.x * ROWS;
  for (int i = threadIdx.x; i < (ROWS + 2) * 66; i += blockDim.x) {
    const int r = i / 66, c = i % 66;
    const int h = h_base + r - 1, w = c - 1;
    float v = 0.0f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = x[(static_cast<size_t>(n) * H + h) * W + w];
    s_in[r][c] = v;
  }
  const int cg = threadIdx.x & 7;   // channels cg*8 .. cg*8+7
  const int pl = threadIdx.x >> 3;  // 0..31
  float wr[8][9], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[cg * 8 + j];
    sh[j] = shift[cg * 8 + j];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = w9[(cg * 8 + j) * 9 + t];
  }
  __syncthreads();
  for (int pix = pl; pix < ROWS * 64; pix += 32) {
    const int r = pix >> 6, c = pix & 63;
    const int h = h_base + r;
    if (h >= H) break;
    float in[9];
#pragma unroll
    for (int dr = 0; dr < 3; ++dr)
#pragma unroll
      for (int dc = 0; dc < 3; ++dc) in[dr * 3 + dc] = s_in[r + dr][c + dc];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.0f;
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(in[t], wr[j][t], a);
      v[j] = fmaxf(fmaf(a, sc[j], sh[j]), 0.0f);
    }
    uint4 q;
    q.x = Elem16<T>::pack2(v[0], v[1]);
    q.y = Elem16<T>::pack2(v[2], v[3]);
    q.z = Elem16<T>::pack2(v[4], v[5]);
    q.w = Elem16<T>::pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * H + h) * W + c) * 64 + cg * 8) = q;
  }
}

int conv_first_launch(const float* x, int NB, int H, int W, const float* w9, const float* scale, const float* shift,
                      void* out, int dtype, cudaStream_t stream) {
  if (W != 64 || NB <= 0 || H <= 0) {
    set_error("conv_first: W must be 64 (got %d)", W);
    return SED_ERR_BAD_SHAPE;
  }
  dim3 grid((H + 7) / 8, NB);
  if (dtype == 0)
    conv_first_kernel<__half><<<grid, 256, 0, stream>>>(x, H, W, w9, scale, shift, reinterpret_cast<__half*>(out));
  else
    conv_first_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x, H, W, w9, scale, shift,
                                                               reinterpret_cast<__nv_bfloat16*>(out));
  return cudaGetLastError() == cudaSuccess ? SED_OK : SED_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 conv / linear launch
// ---------------------------------------------------------------------------------------------
template <typename T, int CIN, int BN, int NT, bool BRES, bool PATCH, int EPI, int SA, int SB>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<CIN, BN, NT, BRES, PATCH, EPI, SA, SB>;
  auto kern = conv_umma_kernel<T, CIN, BN, NT, BRES, PATCH, EPI, SA, SB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  const int groups = (p.num_tiles + NT - 1) / NT;
  int grid;
  if (BRES) {
    int per_slice = num_sms() / p.nslices;
    if (per_slice > groups) per_slice = groups;
    if (per_slice < 1) per_slice = 1;
    grid = per_slice * p.nslices;
  } else {
    const long items = static_cast<long>(groups) * p.nslices;
    grid = items < num_sms() ? static_cast<int>(items) : num_sms();
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("conv_umma launch: %s", cudaGetErrorString(e));
    return SED_ERR_CUDA;
  }
  return SED_OK;
}

template <typename T>
static int conv3x3_dispatch(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvParams& p, int cin, int cout, int mode,
                            int variant, cudaStream_t stream) {
  // (cin, cout, mode) are the seven tensor-core layers of Cnn_9layers.
#define SED_CASE(CIN_, COUT_, MODE_, BN_, NT_, BRES_, SA_P, SB_P, SA_T, SB_T)                                     \
  if (cin == CIN_ && cout == COUT_ && mode == MODE_) {                                                             \
    p.nslices = COUT_ / BN_;                                                                                       \
    if (variant == 0) return launch_cfg<T, CIN_, BN_, NT_, BRES_, true, MODE_, SA_P, SB_P>(tmA, tmB, p, stream);   \
    return launch_cfg<T, CIN_, BN_, NT_, BRES_, false, MODE_, SA_T, SB_T>(tmA, tmB, p, stream);                    \
  }
  //        cin cout mode           BN  NT bres  SA/SB patch  SA/SB tap
  SED_CASE(64, 64, EPI_POOL, 64, 1, true, 4, 1, 6, 1)       // conv_block1.conv2 : weights resident (72 KB)
  SED_CASE(64, 128, EPI_STORE, 128, 1, true, 3, 1, 4, 1)    // conv_block2.conv1 : weights resident (144 KB)
  SED_CASE(128, 128, EPI_POOL, 64, 1, true, 3, 1, 4, 1)     // conv_block2.conv2 : 2 Cout slices resident
  SED_CASE(128, 256, EPI_STORE, 64, 1, true, 3, 1, 4, 1)    // conv_block3.conv1 : 4 Cout slices resident
  SED_CASE(256, 256, EPI_POOL, 256, 2, false, 2, 3, 3, 3)   // conv_block3.conv2 : weights streamed, 2 tiles/CTA
  SED_CASE(256, 512, EPI_STORE, 256, 2, false, 2, 3, 3, 3)  // conv_block4.conv1
  SED_CASE(512, 512, EPI_FREQMEAN, 256, 2, false, 2, 3, 3, 3)  // conv_block4.conv2 (+ freq mean, models.py:668)
#undef SED_CASE
  set_error("conv3x3: unsupported layer (cin=%d, cout=%d, mode=%d)", cin, cout, mode);
  return SED_ERR_UNSUPPORTED;
}

int conv3x3_launch(const void* x, int NB, int H, int W, int cin, const void* wpacked, const float* scale,
                   const float* shift, int cout, int mode, void* out, int dtype, int variant, cudaStream_t stream) {
  if (NB <= 0 || H <= 0 || W <= 0 || (W % 8) != 0 || (cin % 64) != 0) {
    set_error("conv3x3: bad shape NB=%d H=%d W=%d cin=%d", NB, H, W, cin);
    return SED_ERR_BAD_SHAPE;
  }
  if (mode == EPI_FREQMEAN && W != 8) {
    set_error("conv3x3: freq-mean epilogue needs W == 8 (got %d)", W);
    return SED_ERR_BAD_SHAPE;
  }
  if (mode == EPI_POOL && (W % 16) != 0) {
    set_error("conv3x3: pooling epilogue needs W %% 16 == 0 (got %d)", W);
    return SED_ERR_BAD_SHAPE;
  }
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    const uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
    const uint32_t box_patch[4] = {64, 10, 18, 1};
    const uint32_t box_tap[4] = {64, 8, 16, 1};
    int rc = make_map(&tmA, dtype, 4, const_cast<void*>(x), dims, str, variant == 0 ? box_patch : box_tap);
    if (rc) return rc;
  }
  int bn = 0;
  if (cin == 64 && cout == 64) bn = 64;
  else if (cin == 64 && cout == 128) bn = 128;
  else if (cin == 128) bn = 64;
  else bn = 256;
  {
    const uint64_t dims[2] = {(uint64_t)9 * cin, (uint64_t)cout};
    const uint64_t str[1] = {(uint64_t)9 * cin * 2};
    const uint32_t box[2] = {64, (uint32_t)bn};
    int rc = make_map(&tmB, dtype, 2, const_cast<void*>(wpacked), dims, str, box);
    if (rc) return rc;
  }
  ConvParams p{};
  p.NB = NB; p.H = H; p.W = W;
  p.tiles_h = (H + 15) / 16;
  p.tiles_w = W / 8;
  p.num_tiles = NB * p.tiles_h * p.tiles_w;
  p.cout = cout;
  p.scale = scale; p.shift = shift;
  p.out = out; p.out2 = nullptr;
  p.M = 0; p.ldc = 0; p.relu = 1;
  if (dtype == 0) return conv3x3_dispatch<__half>(tmA, tmB, p, cin, cout, mode, variant, stream);
  if (dtype == 1) return conv3x3_dispatch<__nv_bfloat16>(tmA, tmB, p, cin, cout, mode, variant, stream);
  set_error("conv3x3: dtype must be 0 (fp16) or 1 (bf16)");
  return SED_ERR_UNSUPPORTED;
}

int linear_launch(const void* a16, long M, int K, const void* w16, const float* bias, int N, int relu, float* out,
                  void* out16, int dtype, cudaStream_t stream) {
  if (M <= 0 || (N % 128) != 0 || N > 512 * 8) {
    set_error("linear: bad shape M=%ld N=%d K=%d", M, N, K);
    return SED_ERR_BAD_SHAPE;
  }
  if (K != 512 && K != 256) {
    set_error("linear: K must be 256 or 512 (got %d)", K);
    return SED_ERR_UNSUPPORTED;
  }
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, 128};
    int rc = make_map(&tmA, dtype, 2, const_cast<void*>(a16), dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, 128};
    int rc = make_map(&tmB, dtype, 2, const_cast<void*>(w16), dims, str, box);
    if (rc) return rc;
  }
  // the epilogue caches per-channel shift for at most 512 channels: run N in column panels of <= 512
  int rc = SED_OK;
  for (int n0 = 0; n0 < N && rc == SED_OK; n0 += 512) {
    const int npanel = (N - n0) < 512 ? (N - n0) : 512;
    CUtensorMap tmBp = tmB;
    if (n0 != 0) {
      const uint64_t dims[2] = {(uint64_t)K, (uint64_t)npanel};
      const uint64_t str[1] = {(uint64_t)K * 2};
      const uint32_t box[2] = {64, 128};
      rc = make_map(&tmBp, dtype, 2, (char*)const_cast<void*>(w16) + (size_t)n0 * K * 2, dims, str, box);
      if (rc) return rc;
    }
    ConvParams p{};
    p.num_tiles = (int)((M + 127) / 128);
    p.cout = npanel;
    p.nslices = npanel / 128;
    p.scale = nullptr;
    p.shift = bias ? bias + n0 : nullptr;
    p.out = out + n0;
    p.out2 = out16 ? (void*)((char*)out16 + (size_t)n0 * 2) : nullptr;
    p.M = (int)M; p.ldc = N; p.relu = relu;
    if (K == 512) {
      rc = dtype == 0 ? launch_cfg<__half, 512, 128, 1, false, false, EPI_LINEAR, 4, 4>(tmA, tmBp, p, stream)
                      : launch_cfg<__nv_bfloat16, 512, 128, 1, false, false, EPI_LINEAR, 4, 4>(tmA, tmBp, p, stream);
    } else {
      rc = dtype == 0 ? launch_cfg<__half, 256, 128, 1, false, false, EPI_LINEAR, 4, 4>(tmA, tmBp, p, stream)
                      : launch_cfg<__nv_bfloat16, 256, 128, 1, false, false, EPI_LINEAR, 4, 4>(tmA, tmBp, p, stream);
    }
  }
  return rc;
}

}  // namespace sed
