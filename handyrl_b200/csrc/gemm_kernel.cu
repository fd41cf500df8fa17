// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, sm_100a) for the dense contractions of the user's net.
//
//   C[M x N] = A_op[M x K] * B_op[N x K]^T  (+ bias[N])           all fp32 in global memory
//
// north_star: "tensor cores only for the model's Linear/Conv layers where they are dense contractions".  The nets of
// the reference's board games (handyrl/envs/tictactoe.py:52-69, geister.py:101-167) convolve over boards of a few cells;
// fastnet.py runs such a layer as ONE dense matrix product per direction (forward, input gradient, weight gradient),
// which cuBLAS executes as SIMT SGEMM because the learner's contract is fp32 (1e-5 of the reference).  Here the same
// product runs on tcgen05.mma kind::tf32 with the 3xTF32 split
//       a = a_hi + a_lo,  a_hi = a with the 13 low mantissa bits cleared (exactly a TF32 number), a_lo = a - a_hi (exact)
//       a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi                     (dropped: a_lo*b_lo ~ 2^-22 |a||b|)
// accumulated in fp32 in tensor memory: fp32-class accuracy (relative error ~1e-6 of |a||b| sums, tests/test_gemm_gpu.py)
// at tensor-core speed.
//
// Structure (one CTA = one 128-row tile of C x up to 288 columns x one slice of K):
//   * all 16 warps are producers: global fp32 -> registers -> (hi, lo) split -> shared memory in the UMMA canonical
//     K-major SWIZZLE_128B layout (one 128-byte row per operand row and chunk, 16-byte slots XOR-swizzled by the row:
//     conflict-free 16-byte stores both for row-contiguous and for transposing loads); either operand may be stored
//     with its reduction dimension contiguous ("k-major") or strided (transposed on the fly: the weight-gradient
//     product reduces over samples);
//   * two shared-memory stages of 32 reduction elements; one elected thread issues 3 x 4 (x 2 column halves when
//     N > 256) tcgen05.mma per stage and commits them to the stage's mbarrier, which the producers wait on before
//     refilling the stage (TMA is not used: the operands need the hi/lo split on their way in);
//   * the accumulator tile (128 lanes x N columns, fp32) lives in TMEM; the epilogue reads it with tcgen05.ld
//     (one lane quarter per warp), adds the bias and stores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace hrl {

constexpr int kGemmThreads = 512;
constexpr int kTileM = 128;
constexpr int kMaxN = 288;          // columns of one CTA tile (TMEM: 512 fp32 columns; shared memory: 2 stages)
constexpr int kChunkK = 32;         // reduction elements per shared-memory stage
constexpr int kStages = 2;

struct GemmParams {
    const float *A, *B, *bias;
    float *C;
    long long lda, ldb, ldc;
    long long c_split_stride;       // elements between the partial outputs of consecutive K slices
    int M, N, K;
    int a_kmajor, b_kmajor;         // 1: element (r,k) at r*ld + k ; 0: at k*ld + r
    int chunks_per_split;
    int debug;                      // profiling only: 1 = no MMAs, 2 = no loads/stores
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 |
// SBO>>4 <<32 | version 1 <<46 | layout_type SWIZZLE_128B = 2 <<61).  A row of the tile is the 128 bytes (32 reduction
// elements) of one chunk; 8-row groups are 1024 bytes apart (SBO); inside a group the 16-byte slot j of row r sits at
// slot j ^ (r & 7) (Swizzle<3,4,3>).  LBO is not used by swizzled K-major layouts (set to 1).  The k-step inside the
// chunk is selected by advancing the start address by 32 bytes.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor of tcgen05.mma kind::tf32 (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        :
        : "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void split_tf32(const float4 v, float4 &hi, float4 &lo) {
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    lo.x = v.x - hi.x;
    lo.y = v.y - hi.y;
    lo.z = v.z - hi.z;
    lo.w = v.w - hi.w;
}

// ---- operand loaders.  Every thread owns a fixed set of "items" (one row x 4 consecutive reduction elements = one 16-byte
// shared-memory slot per split half); their coordinates are computed once, each chunk only advances the pointers.
struct Item {
    const float *ptr;      // first of the 4 elements in chunk 0 (valid rows only)
    uint32_t slot;         // byte offset of the 16-byte slot inside an operand half: row * 128 + ((j ^ (row & 7)) << 4)
    int k;                 // 4 * j: offset of the quad inside a chunk
    bool live;             // the row exists
};

template <bool KMAJOR>
__device__ __forceinline__ Item make_item(int i, int n_items, const float *base, long long ld, int rows_pad, int rows) {
    Item it;
    int row, j;
    if (KMAJOR) {           // 8 consecutive lanes = the 128 contiguous bytes of one row's chunk: one cache line per quarter
        row = i >> 3;       // warp in global memory, and (swizzle) 8 distinct 16-byte slots in shared memory
        j = i & 7;
    } else {                // consecutive lanes = consecutive rows: coalesced along the contiguous dimension, and the swizzle
        j = i / rows_pad;   // spreads 8 consecutive rows of one slot column over 8 distinct slots
        row = i - j * rows_pad;
    }
    it.live = i < n_items && row < rows;
    it.k = 4 * j;
    it.slot = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
    it.ptr = base + (KMAJOR ? (long long)row * ld + 4 * j : (long long)(4 * j) * ld + row);
    if (i >= n_items) it.slot = 0xFFFFFFFFu;
    return it;
}

template <bool KMAJOR>
__device__ __forceinline__ float4 load_item(const Item &it, long long ld, bool vec, long long advance, int k_left) {
    // k_left = reduction elements from this chunk's start to the end of the operand (>= 32 in every chunk but the last)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!it.live) return v;
    const float *q = it.ptr + advance;
    if (k_left >= kChunkK) {                     // interior chunk: no per-element bounds
        if (KMAJOR) {
            if (vec) return __ldg(reinterpret_cast<const float4 *>(q));
            v.x = __ldg(q); v.y = __ldg(q + 1); v.z = __ldg(q + 2); v.w = __ldg(q + 3);
        } else {
            v.x = __ldg(q); v.y = __ldg(q + ld); v.z = __ldg(q + 2 * ld); v.w = __ldg(q + 3 * ld);
        }
        return v;
    }
    const long long st = KMAJOR ? 1 : ld;          // last, partial chunk
    if (it.k + 0 < k_left) v.x = __ldg(q);
    if (it.k + 1 < k_left) v.y = __ldg(q + st);
    if (it.k + 2 < k_left) v.z = __ldg(q + 2 * st);
    if (it.k + 3 < k_left) v.w = __ldg(q + 3 * st);
    return v;
}

template <bool A_K, bool B_K, int ITEMS_A, int ITEMS_B>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tf32x3_kernel(const GemmParams p, const int n_pad) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // swizzle atoms need 1024-byte alignment
    __shared__ __align__(8) uint64_t bars[kStages + 1];
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * kTileM;
    const int n0 = blockIdx.y * kMaxN;
    const int split = blockIdx.z;
    const int n_here = min(kMaxN, p.N - n0);
    const int rows_a = min(kTileM, p.M - m0);
    const int total_chunks = (p.K + kChunkK - 1) / kChunkK;
    const int c_begin = split * p.chunks_per_split;
    const int c_end = min(total_chunks, c_begin + p.chunks_per_split);

    // stage layout: [A_hi | A_lo | B_hi | B_lo], each [rows][128 B] with the 16-byte slots of a row swizzled
    const uint32_t a_bytes = kTileM * kChunkK * 4, b_bytes = (uint32_t)n_pad * kChunkK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const uint32_t smem_base = smem_u32(smem);

    if (tid == 0) {
        for (int s = 0; s <= kStages; s++) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    const float *Ag = p.A + (A_K ? (long long)m0 * p.lda : (long long)m0);
    const float *Bg = p.B + (B_K ? (long long)n0 * p.ldb : (long long)n0);
    const bool vec_a = A_K && (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool vec_b = B_K && (p.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);
    const int halves = n_pad > 256 ? 2 : 1;
    const int n_mma = n_pad / halves;
    const uint32_t idesc = umma_idesc_tf32(kTileM, n_mma);

    Item ia[ITEMS_A], ib[ITEMS_B];
#pragma unroll
    for (int u = 0; u < ITEMS_A; u++) ia[u] = make_item<A_K>(tid + u * kGemmThreads, kTileM * 8, Ag, p.lda, kTileM, rows_a);
#pragma unroll
    for (int u = 0; u < ITEMS_B; u++) ib[u] = make_item<B_K>(tid + u * kGemmThreads, n_pad * 8, Bg, p.ldb, n_pad, n_here);

    for (int c = c_begin; c < c_end; c++) {
        const int it = c - c_begin, s = it & 1;
        const int k0 = c * kChunkK;
        const int k_left = p.K - k0;
        const long long adv_a = A_K ? (long long)k0 : (long long)k0 * p.lda;
        const long long adv_b = B_K ? (long long)k0 : (long long)k0 * p.ldb;
        // ---- global loads of this chunk (issued before waiting for the stage: latency overlaps the running MMAs)
        float4 va[ITEMS_A], vb[ITEMS_B];
        if (p.debug != 2) {
#pragma unroll
            for (int u = 0; u < ITEMS_A; u++) va[u] = load_item<A_K>(ia[u], p.lda, vec_a, adv_a, k_left);
#pragma unroll
            for (int u = 0; u < ITEMS_B; u++) vb[u] = load_item<B_K>(ib[u], p.ldb, vec_b, adv_b, k_left);
        }
        if (it >= kStages) mbar_wait(smem_u32(&bars[s]), ((it >> 1) - 1) & 1);      // the MMAs that read this stage are done
        if (p.debug != 2) {
            uint8_t *stp = smem + s * stage_bytes;
#pragma unroll
            for (int u = 0; u < ITEMS_A; u++) {
                float4 hi, lo;
                split_tf32(va[u], hi, lo);
                *reinterpret_cast<float4 *>(stp + ia[u].slot) = hi;
                *reinterpret_cast<float4 *>(stp + a_bytes + ia[u].slot) = lo;
            }
#pragma unroll
            for (int u = 0; u < ITEMS_B; u++) {
                if (ib[u].slot != 0xFFFFFFFFu) {
                    float4 hi, lo;
                    split_tf32(vb[u], hi, lo);
                    *reinterpret_cast<float4 *>(stp + 2 * a_bytes + ib[u].slot) = hi;
                    *reinterpret_cast<float4 *>(stp + 2 * a_bytes + b_bytes + ib[u].slot) = lo;
                }
            }
        }
        const uint32_t st = smem_base + s * stage_bytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0 && p.debug == 1) {
            umma_commit(smem_u32(&bars[s]));
            if (c == c_end - 1) umma_commit(smem_u32(&bars[kStages]));
        }
        if (tid == 0 && p.debug != 1) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = st, a_lo = st + a_bytes, b_hi = st + 2 * a_bytes, b_lo = st + 2 * a_bytes + b_bytes;
#pragma unroll
            for (int ks = 0; ks < kChunkK / 8; ks++) {
                for (int h = 0; h < halves; h++) {
                    const uint32_t boff = 32 * ks + h * n_mma * 128;      // n_mma % 8 == 0: whole swizzle atoms
                    const uint32_t aoff = 32 * ks;
                    const uint32_t d = tmem_base + h * n_mma;
                    const uint32_t first = (it == 0 && ks == 0) ? 0u : 1u;
                    // small terms first: a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
                    umma_tf32(d, umma_desc(a_lo + aoff), umma_desc(b_hi + boff), idesc, first);
                    umma_tf32(d, umma_desc(a_hi + aoff), umma_desc(b_lo + boff), idesc, 1u);
                    umma_tf32(d, umma_desc(a_hi + aoff), umma_desc(b_hi + boff), idesc, 1u);
                }
            }
            umma_commit(smem_u32(&bars[s]));
            if (c == c_end - 1) umma_commit(smem_u32(&bars[kStages]));
        }
    }

    // ---- epilogue: TMEM -> registers -> shared-memory tile (padded rows) -> coalesced global stores.
    // (a thread holds ONE row of the accumulator: storing from registers would touch 32 cache lines per warp instruction)
    const int n_chunks_here = c_end - c_begin;
    float *Cg = p.C + (long long)split * p.c_split_stride;
    const int q = warp & 3, group = warp >> 2;
    constexpr int kGroups = kGemmThreads / 128;
    const int cols_per_group = ((n_pad + kGroups - 1) / kGroups + 7) / 8 * 8;
    const int ldt = n_pad + 4;                           // row stride = 16 (mod 128) bytes: conflict-free 16-byte stores
    float *tile = reinterpret_cast<float *>(smem);       // the stages are free once the last MMAs have completed
    if (n_chunks_here > 0) {
        mbar_wait(smem_u32(&bars[kStages]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    {
        float *trow = tile + (q * 32 + (tid & 31)) * ldt;
        for (int cb = 0; cb < cols_per_group; cb += 8) {
            const int col = group * cols_per_group + cb;
            if (col >= n_pad) break;
            uint32_t r[8];
            if (n_chunks_here > 0) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                             : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int e = 0; e < 8; e++) r[e] = 0u;
            }
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                v[e] = __uint_as_float(r[e]);
                if (p.bias != nullptr && col + e < n_here) v[e] += __ldg(p.bias + n0 + col + e);
            }
            reinterpret_cast<float4 *>(trow + col)[0] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4 *>(trow + col)[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();
    {
        const bool vec_c = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cg) & 15) == 0) && (n0 % 4 == 0) && (n_here % 4 == 0);
        const int lane = tid & 31;
        for (int r = warp; r < rows_a; r += kGemmThreads / 32) {
            const float *src = tile + r * ldt;
            float *dst = Cg + (long long)(m0 + r) * p.ldc + n0;
            if (vec_c) {
                for (int c4 = lane; c4 < n_here / 4; c4 += 32) reinterpret_cast<float4 *>(dst)[c4] = reinterpret_cast<const float4 *>(src)[c4];
            } else {
                for (int c1 = lane; c1 < n_here; c1 += 32) dst[c1] = src[c1];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// fixed-order sum of the K-slice partials: out[i] = sum_s partials[s][i]  (deterministic)
__global__ void sum_partials_kernel(const float *__restrict__ partials, int splits, long long n, long long stride, float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = partials[i];
        for (int k = 1; k < splits; k++) s += partials[(long long)k * stride + i];
        out[i] = s;
    }
}

}  // namespace hrl

static int g_gemm_debug = 0;
extern "C" void hrl_gemm_set_debug(int v) { g_gemm_debug = v; }

extern "C" size_t hrl_gemm_workspace_floats(int64_t M, int64_t N, int64_t K, int32_t splits) {
    (void)K;
    return splits > 1 ? (size_t)splits * (size_t)M * (size_t)N : 0;
}

// how many K slices a request for `splits` really produces (whole 32-element chunks per slice, no empty slice)
extern "C" int32_t hrl_gemm_effective_splits(int64_t K, int32_t splits) {
    const int total_chunks = (int)((K + hrl::kChunkK - 1) / hrl::kChunkK);
    if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    const int per = (total_chunks + splits - 1) / splits;
    return (total_chunks + per - 1) / per;
}

extern "C" int hrl_gemm_tf32x3(const float *A, int64_t lda, int32_t a_kmajor, const float *B, int64_t ldb, int32_t b_kmajor,
                               const float *bias, float *C, int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t splits,
                               float *workspace, void *stream_) {
    using namespace hrl;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HRL_REQUIRE(A && B && (C || (splits > 1 && workspace)), HRL_ERR_BAD_ARG, "hrl_gemm_tf32x3: NULL pointer");
    HRL_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), HRL_ERR_BAD_ARG,
                "hrl_gemm_tf32x3: bad dimensions (M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
    HRL_REQUIRE(lda >= (a_kmajor ? K : M) && ldb >= (b_kmajor ? K : N) && ldc >= N, HRL_ERR_BAD_ARG,
                "hrl_gemm_tf32x3: leading dimension smaller than the row length");
    const int total_chunks = (int)((K + kChunkK - 1) / kChunkK);
    if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    HRL_REQUIRE(splits == 1 || (workspace != nullptr && bias == nullptr), HRL_ERR_WORKSPACE,
                "hrl_gemm_tf32x3: a split-K product needs a workspace of hrl_gemm_workspace_floats() floats and no bias");
    const int n_tiles = (int)((N + kMaxN - 1) / kMaxN);
    const int n_widest = (int)(N < kMaxN ? N : kMaxN);
    int n_pad = (n_widest + 15) / 16 * 16;
    if (n_pad > 256) n_pad = (n_pad + 31) / 32 * 32;

    GemmParams p;
    p.A = A; p.B = B; p.bias = bias;
    p.lda = lda; p.ldb = ldb;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.a_kmajor = a_kmajor ? 1 : 0; p.b_kmajor = b_kmajor ? 1 : 0;
    p.chunks_per_split = (total_chunks + splits - 1) / splits;
    p.debug = g_gemm_debug;
    splits = (total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;      // no empty slices
    if (splits > 1) {
        p.C = workspace; p.ldc = N; p.c_split_stride = M * N;
    } else {
        p.C = C; p.ldc = ldc; p.c_split_stride = 0;
    }
    const size_t smem_bytes = 1024 + (size_t)kStages * (2 * (size_t)kTileM * kChunkK * 4 + 2 * (size_t)n_pad * kChunkK * 4);
    const dim3 grid((unsigned)((M + kTileM - 1) / kTileM), (unsigned)n_tiles, (unsigned)splits);
    const int items_b = (n_pad * 8 + kGemmThreads - 1) / kGemmThreads;
    constexpr int IA = kTileM * 8 / kGemmThreads;
#define HRL_GEMM_LAUNCH2(AK, BK, IB)                                                                                       \
    {                                                                                                                     \
        HRL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<AK, BK, IA, IB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            (int)smem_bytes));                                                            \
        gemm_tf32x3_kernel<AK, BK, IA, IB><<<grid, kGemmThreads, smem_bytes, stream>>>(p, n_pad);                          \
    }
#define HRL_GEMM_LAUNCH(IB)                                                                    \
    {                                                                                         \
        if (p.a_kmajor && p.b_kmajor) HRL_GEMM_LAUNCH2(true, true, IB)                        \
        else if (p.a_kmajor) HRL_GEMM_LAUNCH2(true, false, IB)                                \
        else if (p.b_kmajor) HRL_GEMM_LAUNCH2(false, true, IB)                                \
        else HRL_GEMM_LAUNCH2(false, false, IB)                                               \
    }
    if (items_b <= 1) HRL_GEMM_LAUNCH(1)
    else if (items_b <= 3) HRL_GEMM_LAUNCH(3)
    else HRL_GEMM_LAUNCH(5)
#undef HRL_GEMM_LAUNCH
#undef HRL_GEMM_LAUNCH2
    HRL_CUDA_CHECK(cudaGetLastError());
    if (splits > 1 && C != nullptr) {       // C == NULL: the caller consumes the slice partials itself (hrl_board_fold)
        const long long n = (long long)M * N;
        HRL_REQUIRE(ldc == N, HRL_ERR_UNSUPPORTED, "hrl_gemm_tf32x3: split-K output must be dense (ldc == N)");
        int blocks = (int)((n + 255) / 256);
        if (blocks > 1184) blocks = 1184;
        sum_partials_kernel<<<blocks, 256, 0, stream>>>(workspace, splits, n, n, C);
        HRL_CUDA_CHECK(cudaGetLastError());
    }
    return HRL_OK;
}
