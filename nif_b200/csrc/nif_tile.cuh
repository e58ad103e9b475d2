// Tile configuration and the register-tile micro-kernel shared by the forward and reverse passes.
#pragma once
#include "nif_common.cuh"

// A CTA owns TB rows (points).  Every thread owns MP = 4*GP rows x MJ = 4*GJ columns of the
// current layer's output.  Activations live in shared memory "k-major":  act[i][row], XOR-swizzled
// in 16-byte units so that both the broadcast reads of the micro-kernel and the transposed
// write-back of a layer's output stay (almost) conflict free.
template <int NP_, int TB_, int GP_, int GJ_>
struct TileCfg {
  static constexpr int NP = NP_, TB = TB_, GP = GP_, GJ = GJ_;
  static constexpr int MP = 4 * GP, MJ = 4 * GJ;
  static constexpr int TX = TB / MP, TY = NP / MJ, NT = TX * TY;
  static constexpr int NIS = NP < 64 ? NP : 64;  // matrix rows per pipeline stage
  static constexpr int NH = NP / NIS;            // stages per (layer, kappa)
  static constexpr int STAGE_FLOATS = NIS * NP;
  static constexpr int PSTR = TB / GP;  // row distance between a thread's row groups
  static constexpr int JSTR = NP / GJ;  // column distance between a thread's column groups
  static_assert(TB % 32 == 0 && NT % 32 == 0, "tile shape");
  static_assert(TY <= 32 && (32 % TY) == 0, "a row group must sit inside one warp");
};

using Cfg32 = TileCfg<32, 128, 1, 1>;   // 32 x 8  threads
using Cfg64 = TileCfg<64, 128, 2, 1>;   // 16 x 16 threads
using Cfg128 = TileCfg<128, 64, 1, 2>;  // 16 x 16 threads

template <class C>
__device__ __forceinline__ int act_idx(int i, int p) {
  return i * C::TB + (p ^ (((i >> 2) & 7) << 2));
}

// thread -> rows / columns
template <class C>
__device__ __forceinline__ int row_of(int tp, int r) {  // r in [0, MP)
  return (r >> 2) * C::PSTR + tp * 4 + (r & 3);
}
template <class C>
__device__ __forceinline__ int col_of(int tj, int c) {  // c in [0, MJ)
  return (c >> 2) * C::JSTR + tj * 4 + (c & 3);
}

// t[r][c] += sum_{ii < NIS} A[i_base + ii][row r] * Bst[ii][col c]
template <class C>
__device__ __forceinline__ void mk_gemm(const float* __restrict__ A, const float* __restrict__ Bst, int i_base,
                                        float (&t)[C::MP][C::MJ], int tp, int tj) {
#pragma unroll 2
  for (int i4 = 0; i4 < C::NIS / 4; ++i4) {
    const int ig = i_base + i4 * 4;
    const int sw = ((ig >> 2) & 7) << 2;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float4 a[C::GP], b[C::GJ];
#pragma unroll
      for (int gp = 0; gp < C::GP; ++gp)
        a[gp] = *reinterpret_cast<const float4*>(&A[(ig + u) * C::TB + ((gp * C::PSTR + tp * 4) ^ sw)]);
#pragma unroll
      for (int gj = 0; gj < C::GJ; ++gj)
        b[gj] = *reinterpret_cast<const float4*>(&Bst[(i4 * 4 + u) * C::NP + gj * C::JSTR + tj * 4]);
#pragma unroll
      for (int gp = 0; gp < C::GP; ++gp) {
        const float av[4] = {a[gp].x, a[gp].y, a[gp].z, a[gp].w};
#pragma unroll
        for (int gj = 0; gj < C::GJ; ++gj) {
          const float bv[4] = {b[gj].x, b[gj].y, b[gj].z, b[gj].w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) t[gp * 4 + e][gj * 4 + f] = fmaf(av[e], bv[f], t[gp * 4 + e][gj * 4 + f]);
        }
      }
    }
  }
}

// Weight-stream pipeline: 2 stages, filled by one elected thread with cp.async.bulk and an
// mbarrier per stage.  The chunk sequence of a CTA is  tile -> hidden layer -> kappa -> half.
template <class C>
struct WeightStream {
  float* stage;    // [2][STAGE_FLOATS]
  uint64_t* bar;   // [2]
  long long total, issued, consumed;
  int chunks_per_tile, H, K1;
  const float* packed;      // base of all packed images
  long long packed_floats;  // image stride
  long long sec_off;        // off_MH (forward) or off_MHT (reverse)
  long long tiles_per_group;
  bool reverse;             // reverse pass walks hidden layers H-1 .. 0

  __device__ __forceinline__ void issue_one() {
    const long long c = issued;
    const long long t = c / chunks_per_tile;
    int r = static_cast<int>(c - t * chunks_per_tile);
    const long long tile = blockIdx.x + t * gridDim.x;
    const long long g = tile / tiles_per_group;
    int h = r / (K1 * C::NH);
    r -= h * (K1 * C::NH);
    if (reverse) h = H - 1 - h;
    const int kappa = r / C::NH, hf = r - kappa * C::NH;
    const float* src = packed + g * packed_floats + sec_off +
                       ((static_cast<long long>(h) * K1 + kappa) * C::NP + hf * C::NIS) * C::NP;
    const int s = static_cast<int>(c & 1);
    mbar_expect_tx(&bar[s], C::STAGE_FLOATS * 4);
    bulk_g2s(stage + s * C::STAGE_FLOATS, src, C::STAGE_FLOATS * 4, &bar[s]);
    ++issued;
  }
  // all threads: wait for the next chunk, return its stage pointer
  __device__ __forceinline__ const float* acquire() {
    const int s = static_cast<int>(consumed & 1);
    mbar_wait(&bar[s], static_cast<uint32_t>((consumed >> 1) & 1));
    return stage + s * C::STAGE_FLOATS;
  }
  // all threads, after the chunk has been read (contains the CTA barrier)
  __device__ __forceinline__ void release() {
    __syncthreads();
    ++consumed;
    if (threadIdx.x == 0 && issued < total) issue_one();
  }
};
