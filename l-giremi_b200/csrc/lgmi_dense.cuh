// lgmi_dense.cuh -- the deep-unit path: contingency counts as a dense int8
// contraction on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// For a unit with many reads the nine cells of every pair's 3x3 table are the
// nine blocks of a Gram matrix
//        G[a][b][i][j] = sum_r X[a][i][r] * X[b][j][r],   a, b in {other, minor, major}
// of the 0/1 indicator matrix X (3*S_pad rows, K_pad = 32*W columns, one byte per
// read, K-contiguous).  T[a*3+b] of pair (i, j) is G[a][b][i][j]; integer and
// exact (s32 accumulators, counts <= R < 2^31).
//
//   k_expand_planes   bit-planes [M | m | C] -> X rows in label order [O | m | M]
//   k_gram_i8         one 128 x 256 output tile per CTA iteration:
//                       warp 0  TMA producer (cp.async.bulk.tensor, 128B swizzle)
//                       warp 1  single-thread tcgen05.mma.kind::i8 issuer, M=128 N=256 K=32
//                       warp 2  TMEM allocation (256 columns of s32 accumulators)
//                       warps 4-7  epilogue: tcgen05.ld -> global scratch
//                     4-stage smem ring (48 KB per stage) with full/empty mbarriers.
//                     Only tiles that contain some pair i < j are computed.
// k_count / k_pairs_generic then read the tables from the scratch instead of
// popcounting (DevUnit::gram_off).
//
// Reference semantics: the counts of /root/reference/src/giremi/mutual_information.py:15-40
// (labels over the common reads), bit-exact.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lgmi {

constexpr int kDenseBM = 128;      // tile rows   (sites i of plane a)
constexpr int kDenseBN = 256;      // tile columns (sites j of plane b)
constexpr int kDenseBK = 128;      // bytes of K per stage == one 128B swizzle row
constexpr int kDenseStages = 4;
constexpr int kDenseThreads = 256;
constexpr uint32_t kDenseStageBytes = (kDenseBM + kDenseBN) * kDenseBK;  // 48 KB
constexpr uint32_t kDenseSmemBytes = kDenseStages * kDenseStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t kDenseTmemCols = 256;

struct DenseTile {
  uint8_t a, b;       // planes (label order: 0 other, 1 minor, 2 major)
  uint16_t I, J;      // site blocks: rows [128 I, +128), columns [256 J, +256)
  uint16_t partial;   // 1: the tile's K range is shared by several items (last wave): results are added atomically
                      //    into a zeroed tile
  uint32_t kb0, kb1;  // k-blocks [kb0, kb1) of this work item (32 bits: a unit may have millions of reads)
};
static_assert(sizeof(DenseTile) == 16, "DenseTile layout");

struct DenseParams {
  const DenseTile* tiles;
  uint32_t n_tiles;
  uint32_t k_blocks;  // K_pad / 128
  uint32_t S_pad;     // multiple of 256
  uint32_t* gram;     // [9][S_pad][S_pad]
  uint32_t* error;    // set to 1 if a barrier wait ran out (never in a correct run)
};

// ------------------------------------------------------------------ expansion
// 4 bits -> 4 bytes of 0/1
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

// one thread per 16 output bytes of one (plane, site) row
__global__ void __launch_bounds__(256) k_expand_planes(const uint32_t* __restrict__ planes, uint32_t S, uint32_t W,
                                                       uint32_t S_pad, uint8_t* __restrict__ X) {
  const uint64_t K_pad = 32ull * W;
  const uint64_t per_row = K_pad / 16u;  // 16-byte groups per row
  const uint64_t total = 3ull * S * per_row;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = e / per_row;
    const uint32_t g = (uint32_t)(e - row * per_row);  // 16 reads: half of word g/2
    const uint32_t s = (uint32_t)(row % S), label = (uint32_t)(row / S);
    const uint32_t* src = planes + (size_t)s * 3u * W;
    const uint32_t w = g >> 1;
    const uint32_t M = src[w], m = src[W + w], C = src[2u * W + w];
    uint32_t bits = (label == 2u) ? (M & C) : (label == 1u) ? (m & C & ~M) : (C & ~M & ~m);
    bits = (bits >> ((g & 1u) * 16u)) & 0xffffu;
    uint4 out;
    out.x = spread4(bits & 15u);
    out.y = spread4((bits >> 4) & 15u);
    out.z = spread4((bits >> 8) & 15u);
    out.w = spread4(bits >> 12);
    *reinterpret_cast<uint4*>(X + ((size_t)label * S_pad + s) * K_pad + (size_t)g * 16u) = out;
  }
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug becomes an error flag + trap instead of a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t* error) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u && clock64() - t0 > 4000000000LL) {  // ~2 s: no tile takes milliseconds
      if (error) atomicExch(error, 1u);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile whose rows are 128 bytes: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) /* LBO = 16 B (unused with swizzle) */ |
         (64ull << 32) /* SBO = 1024 B */ | (1ull << 46) /* descriptor version: Blackwell */ |
         (2ull << 61) /* SWIZZLE_128B */;
}

// kind::i8 instruction descriptor: D = s32, A = B = unsigned 8 bit, both K-major
__device__ __forceinline__ uint32_t umma_idesc_u8(uint32_t M, uint32_t N) {
  return (2u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrives on the mbarrier when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_load_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// zero the 128 x 256 output tiles that several work items add into (one CTA per listed item; an item whose
// kb0 is 0 zeroes its tile, the others skip)
__global__ void __launch_bounds__(256) k_zero_partial_tiles(const DenseTile* __restrict__ tiles, uint32_t n_tiles,
                                                            uint32_t S_pad, uint32_t* __restrict__ gram) {
  const DenseTile tile = tiles[blockIdx.x];
  if (blockIdx.x >= n_tiles || !tile.partial || tile.kb0 != 0) return;
  uint32_t* base = gram + ((size_t)(tile.a * 3u + tile.b) * S_pad + (size_t)tile.I * kDenseBM) * S_pad + (size_t)tile.J * kDenseBN;
  for (uint32_t e = threadIdx.x; e < (uint32_t)kDenseBM * (kDenseBN / 4); e += blockDim.x) {
    const uint32_t row = e / (kDenseBN / 4), q = e % (kDenseBN / 4);
    reinterpret_cast<uint4*>(base + (size_t)row * S_pad)[q] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------ the GEMM
__global__ void __launch_bounds__(kDenseThreads, 1) k_gram_i8(const __grid_constant__ CUtensorMap tmap,
                                                              const DenseParams P) {
  extern __shared__ uint8_t dense_smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dense_smem_raw) + 1023u) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDenseStages * kDenseStageBytes);
  uint64_t* full = bars;                       // [stages]  TMA -> MMA
  uint64_t* empty = bars + kDenseStages;       // [stages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kDenseStages;   // MMA -> epilogue
  uint64_t* acc_empty = acc_full + 1;             // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int s = 0; s < kDenseStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);  // the four epilogue warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kDenseTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const DenseTile tile = P.tiles[t];
        const int32_t row_a = (int32_t)(tile.a * P.S_pad + tile.I * kDenseBM);
        const int32_t row_b = (int32_t)(tile.b * P.S_pad + tile.J * kDenseBN);
        for (uint32_t kb = tile.kb0; kb < tile.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, P.error);
          uint8_t* sa = smem + stage * kDenseStageBytes;
          uint8_t* sb = sa + kDenseBM * kDenseBK;
          mbar_expect_tx(&full[stage], kDenseStageBytes);
          const int32_t k0 = (int32_t)(kb * kDenseBK);
          tma_load_2d(sa, &tmap, &full[stage], k0, row_a);
          tma_load_2d(sb, &tmap, &full[stage], k0, row_b);
          tma_load_2d(sb + 128 * kDenseBK, &tmap, &full[stage], k0, row_b + 128);
          if (++stage == kDenseStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      const uint32_t idesc = umma_idesc_u8(kDenseBM, kDenseBN);
      uint32_t stage = 0, phase = 0, acc_phase = 0;
      for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const DenseTile tile = P.tiles[t];
        mbar_wait(acc_empty, acc_phase ^ 1u, P.error);  // epilogue has drained the accumulator
        tc_fence_after();
        for (uint32_t kb = tile.kb0; kb < tile.kb1; ++kb) {
          mbar_wait(&full[stage], phase, P.error);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kDenseStageBytes);
          const uint32_t sb = sa + kDenseBM * kDenseBK;
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
#pragma unroll
          for (uint32_t k = 0; k < kDenseBK / 32; ++k)  // K = 32 bytes per instruction: +32 B on both operands
            umma_i8(tmem_base, da + 2ull * k, db + 2ull * k, idesc, ((kb - tile.kb0) | k) != 0u);
          umma_commit(&empty[stage]);  // smem slot free once these MMAs have read it
          if (++stage == kDenseStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(acc_full);  // accumulator complete
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {  // ---------------- epilogue: TMEM -> registers -> global
    const uint32_t q = warp - 4u;  // TMEM lane quarter of this warp
    uint32_t acc_phase = 0;
    for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
      const DenseTile tile = P.tiles[t];
      mbar_wait(acc_full, acc_phase, P.error);
      tc_fence_after();
      const uint32_t row = tile.I * kDenseBM + q * 32u + lane;
      uint32_t* out = P.gram + ((size_t)(tile.a * 3u + tile.b) * P.S_pad + row) * P.S_pad + (size_t)tile.J * kDenseBN;
#pragma unroll 1
      for (uint32_t c = 0; c < kDenseBN / 32; ++c) {
        uint32_t r[32];
        tmem_load_32x32(tmem_base + ((q * 32u) << 16) + c * 32u, r);
        if (tile.partial) {
#pragma unroll
          for (int v = 0; v < 32; ++v) atomicAdd(out + c * 32u + v, r[v]);  // (RED: no return value is used)
        } else {
#pragma unroll
          for (int v = 0; v < 8; ++v)
            reinterpret_cast<uint4*>(out + c * 32u)[v] = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
      acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kDenseTmemCols)
                 : "memory");
  }
}

}  // namespace lgmi
