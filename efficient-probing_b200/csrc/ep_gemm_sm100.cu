// Persistent tcgen05 GEMM (operands through TMA, fp32 accumulate in TMEM) for the small dense contractions of
// the EP head: value projection of the pooled tokens and its two gradients (batched over the M queries), the
// classifier (probe_heads.py:76) and its input gradient.
// Operand types: bf16 (kind::f16; the callers pass fp32 data as bf16 hi/lo pairs) or fp32 read as tf32
// (kind::tf32, K-major only).  fp32-accurate products come from three bf16 terms, either concatenated along K by
// the caller or -- GemmTC::x3 -- issued here from hi and lo tiles that share a pipeline stage.
// Each CTA walks 128 x NT output tiles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
// allocator, warps 4-11 = epilogue (TMEM -> registers -> 256-bit global stores); the accumulator is
// double-buffered in TMEM so a tile's epilogue overlaps the next tile's loads and MMAs.
// Either operand may be K-major (contraction index contiguous in memory) or MN-major (output index
// contiguous); the same 128-byte-swizzled shared-memory bytes serve both through the descriptor's
// major bit, so no operand is ever transposed or converted on the way.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "ep_ptx.cuh"
#include "ep_common.cuh"

namespace ep {
using namespace ptx;

constexpr int kGK = 32;                       // contraction elements per stage: one 128-byte swizzle row (32 fp32 / 64 bf16)
constexpr int kGStages = 4;

struct GemmTC {
  int I, J, K, Z;                             // C is I x J per batch entry z < Z, contraction K
  int NT;                                     // column tile (multiple of 32, <= 256)
  int a_mn, b_mn;                             // operand majors
  int a_swap, b_swap;                         // 0: coords (c0, row/k, z); 1: coords (c0, z, row/k)
  int a_zdiv, b_zdiv;                         // operand batch coordinate = z / zdiv (broadcast over z when huge)
  // x3 (bf16 only): 3-term product of hi/lo operand pairs.  Every stage holds the hi AND lo tile of both operands
  // and issues A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, so each operand byte is fetched once (a [hi|lo|hi] x [hi|hi|lo]
  // concatenation along K would fetch the hi tiles twice).  The lo tile of A sits at batch coordinate + a_lo_z
  // and K coordinate + a_lo_k of the same tensor map, the lo tile of B at K coordinate + b_lo_k.
  // (MN-major operands whose lo copy sits in the same row: output coordinate + a_lo_mn / b_lo_mn instead)
  int x3, a_zmul, a_lo_z, a_lo_k, b_lo_k, a_lo_mn, b_lo_mn;
  float* C; const float* bias;
  long long c_row, c_col, c_z, bias_z;        // element strides of C and bias
  int round_tf32;                             // round the stored result to tf32 (it feeds another TF32 GEMM)
  int stages;                                 // smem ring depth
  int bf16;                                   // operands are bf16 (kind::f16, 64 elements per stage) instead of fp32/tf32
  // epilogue mode 1 (projection backward, C = dP[b, z=m, d]): instead of storing fp32 dP, write it as the
  // bf16 hi/lo operand rows (B, Jrows, D) the dA kernel consumes -- dP never touches memory in fp32
  int epi_mode;
  __nv_bfloat16* hl;                          // (B, Jrows, D)
  int Jrows;
};

__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 256-bit global store (sm_100: STG.E.256): 32-byte aligned address
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(384, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmTC g) {
  // persistent: each CTA walks output tiles (column tile fastest, then row tile, then batch); the smem ring
  // runs on across tiles and the accumulator is double-buffered in TMEM, so the epilogue of one tile
  // overlaps the loads and MMAs of the next and the per-CTA set-up cost is paid once
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = 128u * 128u;                       // 128 rows (or 4 atoms x 32 k-rows) x 128 B
  const uint32_t b_bytes = (uint32_t)g.NT * 128u;
  const uint32_t a_lo_off = a_bytes + b_bytes;                // x3 stage: [A_hi | B_hi | A_lo | B_lo]
  const uint32_t stage_bytes = (g.x3 ? 2u : 1u) * (a_bytes + b_bytes);
  const uint32_t bar_base = smem_base + (uint32_t)g.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (g.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * g.stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * g.stages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * g.stages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tmem_cols = g.NT <= 16 ? 32 : g.NT <= 32 ? 64 : g.NT <= 64 ? 128 : g.NT <= 128 ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 256); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_a); prefetch_tmap(&tm_b); }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                                  // the operands are another kernel's output: from here on
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int gk = g.bf16 ? 2 * kGK : kGK;                      // elements per 128-byte row = K elements per stage
  const int nk = (g.K + gk - 1) / gk;
  // MN-major operand tile of a stage: atoms of [gk k-rows x 128 bytes of mn]; UMMA K = 8 (tf32) / 16 (bf16) rows
  const int mn_elems = gk;                                    // mn elements per 128-byte atom row
  const int mn_atoms = 128 / mn_elems;                        // atoms covering the 128 rows of the A tile
  const uint32_t atom_bytes = (uint32_t)gk * 128u;            // 4 KB (tf32) / 8 KB (bf16)
  const uint32_t mn_adv = g.bf16 ? 2048u : 1024u;             // bytes per UMMA K step along the k-rows
  const int n_ct = (g.J + g.NT - 1) / g.NT, n_rt = (g.I + 127) / 128;
  const int ntiles = n_ct * n_rt * g.Z;
  auto tile_coords = [&](int t, int& i0, int& j0, int& z) {
    const int ct = t % n_ct, r = t / n_ct;
    j0 = ct * g.NT; i0 = (r % n_rt) * 128; z = r / n_rt;
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int i0, j0, z;
        tile_coords(t, i0, j0, z);
        const int za = (z / g.a_zdiv) * g.a_zmul, zb = z / g.b_zdiv;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), stage_bytes);
          const int k0 = kc * gk;
          auto load_pair = [&](uint32_t d, int ka, int zaa, int kb, int ia, int jb) {
            if (!g.a_mn) {                                     // [128 rows x one 128-byte row of k]
              tma_load_3d(d, &tm_a, full_bar(s), ka, g.a_swap ? zaa : i0, g.a_swap ? i0 : zaa);
            } else {                                           // mn atoms of [gk k-rows x 128 B of mn]
              for (int a = 0; a < mn_atoms; ++a)
                tma_load_3d(d + (uint32_t)a * atom_bytes, &tm_a, full_bar(s), ia + mn_elems * a, g.a_swap ? zaa : ka,
                            g.a_swap ? ka : zaa);
            }
            const uint32_t bdst = d + a_bytes;
            if (!g.b_mn) {                                     // [NT rows x one 128-byte row of k]
              tma_load_3d(bdst, &tm_b, full_bar(s), kb, g.b_swap ? zb : j0, g.b_swap ? j0 : zb);
            } else {
              for (int a = 0; a < g.NT / mn_elems; ++a)
                tma_load_3d(bdst + (uint32_t)a * atom_bytes, &tm_b, full_bar(s), jb + mn_elems * a, g.b_swap ? zb : kb,
                            g.b_swap ? kb : zb);
            }
          };
          load_pair(dst, k0, za, k0, i0, j0);
          if (g.x3) load_pair(dst + a_lo_off, k0 + g.a_lo_k, za + g.a_lo_z, k0 + g.b_lo_k, i0 + g.a_lo_mn, j0 + g.b_lo_mn);
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = g.bf16 ? idesc_bf16(128, g.NT, g.a_mn, g.b_mn) : idesc_tf32(128, g.NT, g.a_mn, g.b_mn);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(tempty_bar(buf), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * g.NT);
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t asm_ = smem_base + (uint32_t)s * stage_bytes, bsm = asm_ + a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                        // UMMA K = 8 fp32
            const uint64_t ad = g.a_mn ? smem_desc_sw128(asm_ + mn_adv * k, atom_bytes, 1024)
                                       : smem_desc_sw128(asm_ + 32u * k, 16, 1024);
            const uint64_t bd = g.b_mn ? smem_desc_sw128(bsm + mn_adv * k, atom_bytes, 1024)
                                       : smem_desc_sw128(bsm + 32u * k, 16, 1024);
            if (g.x3) {                                        // small terms first, then hi.hi
              const uint64_t lo_step = (uint64_t)(a_lo_off >> 4);  // descriptor start-address field counts 16 B
              umma_bf16(acc, ad + lo_step, bd, idesc, (uint32_t)((kc | k) != 0));
              umma_bf16(acc, ad, bd + lo_step, idesc, 1u);
              umma_bf16(acc, ad, bd, idesc, 1u);
            } else if (g.bf16) {
              umma_bf16(acc, ad, bd, idesc, (uint32_t)((kc | k) != 0));
            } else {
              umma_tf32(acc, ad, bd, idesc, (uint32_t)((kc | k) != 0));
            }
          }
          umma_commit(empty_bar(s));
          if (++s == g.stages) { s = 0; ph ^= 1u; }
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else if (warp >= 4) {
    const int wq = (warp - 4) & 3, eh = (warp - 4) >> 2;        // lane quadrant, half of the 16-column units
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      int i0, j0, z;
      tile_coords(t, i0, j0, z);
      const int buf = it & 1;
      mbar_wait(tfull_bar(buf), ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
      const int row = i0 + wq * 32 + lane;
      const uint32_t acc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * g.NT);
      if (g.epi_mode == 1) {
        // row = sample b, z = query m, columns = d (J == D): dP leaves TMEM as bf16 hi/lo operand rows
        __nv_bfloat16* hrow = g.hl + ((size_t)row * g.Jrows + 2 * z) * g.J;
        for (int c0 = 16 * eh; c0 < g.NT; c0 += 32) {
          uint32_t r[16];
          tmem_ld16(acc + (uint32_t)c0, r);
          tmem_ld_wait();
          const int col = j0 + c0;
          if (row < g.I && col + 16 <= g.J) {
            // packed conversions (two values per instruction on the FMA pipe; the single-value F2F shares the
            // 16-lane XU pipe)
            __align__(32) uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float v0 = __uint_as_float(r[2 * i]), v1 = __uint_as_float(r[2 * i + 1]);
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
              hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __uint_as_float(hi[i] << 16), v1 - __uint_as_float(hi[i] & 0xffff0000u));
              lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            st_global_256(hrow + col, hi);                                             // one full 32-byte sector
            st_global_256(hrow + g.J + col, lo);                                       // per lane and store
          }
        }
      } else {
        float* crow = g.C + (long long)z * g.c_z + (long long)row * g.c_row;
        const float* bias = g.bias ? g.bias + (long long)z * g.bias_z : nullptr;
        const bool vec = g.c_col == 1 && (g.c_row % 8) == 0 && (g.c_z % 8) == 0 && (j0 % 8) == 0 &&
                         ((reinterpret_cast<uintptr_t>(g.C) & 31) == 0);
        for (int c0 = 16 * eh; c0 < g.NT; c0 += 32) {
          uint32_t r[16];
          tmem_ld16(acc + (uint32_t)c0, r);
          tmem_ld_wait();
          if (row < g.I) {
            if (vec && j0 + c0 + 16 <= g.J) {                  // 64 contiguous bytes per lane
              __align__(32) float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(r[i]) + (bias ? __ldg(bias + j0 + c0 + i) : 0.f);
                if (g.round_tf32) v[i] = to_tf32(v[i]);
              }
              st_global_256(crow + j0 + c0, reinterpret_cast<const uint32_t*>(v));         // full 32-byte sectors
              st_global_256(crow + j0 + c0 + 8, reinterpret_cast<const uint32_t*>(v) + 8);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int col = j0 + c0 + i;
                if (col < g.J) {
                  float v = __uint_as_float(r[i]) + (bias ? __ldg(bias + col) : 0.f);
                  if (g.round_tf32) v = to_tf32(v);
                  crow[(long long)col * g.c_col] = v;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  pdl_trigger();                                               // (at the end: see ep_fused_sm100.cu)
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
}  // namespace

// fp32 / bf16 tensor (d2, d1, d0) with element strides (s2, s1, 1); box (b2, b1, one 128-byte row)
int make_tmap_f32(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                  uint32_t b1, uint32_t b2, int bf16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return EP_ERR_DEVICE;
  const uint64_t es = bf16 ? 2 : 4;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * es, s2 * es};
  cuuint32_t box[3] = {bf16 ? 64u : 32u, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                  const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : EP_ERR_UNSUPPORTED;
}

bool gemm_tc_available() { return encode_fn() != nullptr; }
int launch_gemm_tc(const CUtensorMap& tm_a, const CUtensorMap& tm_b, GemmTC g, int Z, cudaStream_t s);

int launch_gemm_tc(const CUtensorMap& tm_a, const CUtensorMap& tm_b, GemmTC g, int Z, cudaStream_t s) {
  const size_t per_stage = (g.x3 ? 2 : 1) * (128 * 128 + (size_t)g.NT * 128);
  g.stages = (int)std::max<size_t>(2, std::min<size_t>(6, (g.x3 ? 216 * 1024 : 190 * 1024) / per_stage));
  const size_t smem = (size_t)g.stages * per_stage + 1024 + 256;
  if (smem > 224 * 1024) return EP_ERR_UNSUPPORTED;
  EP_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  g.Z = Z;
  const int ntiles = ((g.J + g.NT - 1) / g.NT) * ((g.I + 127) / 128) * Z;
  EP_CUDA(launch_pdl(gemm_tf32_kernel, dim3(std::min(ntiles, kNumSMs)), dim3(384), smem, s, tm_a, tm_b, g));
  EP_LAUNCH_CHECK();
  return 0;
}

static int side_tmap(CUtensorMap* m, const TcSide& sd, int tile_rows) {
  // K-major: box = one 128-byte row of k x tile_rows rows; MN-major: box = 128 bytes of mn x (32 | 64) k-rows
  const uint32_t r = sd.mn_major ? (sd.bf16 ? 64u : 32u) : (uint32_t)tile_rows;
  return make_tmap_f32(m, sd.base, sd.d0, sd.d1, sd.d2, sd.s1, sd.s2, sd.swap ? 1u : r, sd.swap ? r : 1u, sd.bf16);
}

int tc_gemm_dp(const TcSide& A, const TcSide& B, int I, int J, int K, int Z, int NT, void* hl, int Jrows,
               cudaStream_t s) {
  CUtensorMap ta, tb;
  int rc;
  if ((rc = side_tmap(&ta, A, 128))) return rc;
  if ((rc = side_tmap(&tb, B, NT))) return rc;
  GemmTC g{};
  g.I = I; g.J = J; g.K = K; g.NT = NT;
  g.a_mn = A.mn_major; g.b_mn = B.mn_major; g.a_swap = A.swap; g.b_swap = B.swap;
  g.a_zdiv = A.zdiv > 0 ? A.zdiv : 1; g.b_zdiv = B.zdiv > 0 ? B.zdiv : 1;
  g.epi_mode = 1; g.hl = (__nv_bfloat16*)hl; g.Jrows = Jrows;
  g.bf16 = A.bf16; g.a_zmul = 1;
  return launch_gemm_tc(ta, tb, g, Z, s);
}

int tc_gemm(const TcSide& A, const TcSide& B, int I, int J, int K, int Z, int NT, float* C, long long c_row,
            long long c_col, long long c_z, const float* bias, long long bias_z, int round_out, cudaStream_t s) {
  CUtensorMap ta, tb;
  int rc;
  if ((rc = side_tmap(&ta, A, 128))) return rc;
  if ((rc = side_tmap(&tb, B, NT))) return rc;
  GemmTC g{};
  g.I = I; g.J = J; g.K = K; g.NT = NT;
  g.a_mn = A.mn_major; g.b_mn = B.mn_major; g.a_swap = A.swap; g.b_swap = B.swap;
  g.a_zdiv = A.zdiv > 0 ? A.zdiv : 1; g.b_zdiv = B.zdiv > 0 ? B.zdiv : 1;
  g.C = C; g.bias = bias; g.c_row = c_row; g.c_col = c_col; g.c_z = c_z; g.bias_z = bias_z;
  g.round_tf32 = round_out;
  g.bf16 = A.bf16;
  g.a_zmul = 1;
  if (A.lo_k || A.lo_z || B.lo_k || A.lo_mn || B.lo_mn) {                            // hi/lo operand pairs: 3-term product
    if (!A.bf16 || !B.bf16) return EP_ERR_UNSUPPORTED;
    g.x3 = 1; g.a_zmul = A.zmul > 0 ? A.zmul : 1; g.a_lo_z = A.lo_z; g.a_lo_k = A.lo_k; g.b_lo_k = B.lo_k;
    g.a_lo_mn = A.lo_mn; g.b_lo_mn = B.lo_mn;
  }
  return launch_gemm_tc(ta, tb, g, Z, s);
}

}  // namespace ep
