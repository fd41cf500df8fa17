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
#include <string.h>

#include "common.cuh"

// profiling / test hook (include/hrl_b200.h): 1 = no MMAs, 2 = no operand loads, +64 = A operand never staged through shared memory
static int g_gemm_debug = 0;
extern "C" void hrl_gemm_set_debug(int v) { g_gemm_debug = v; }

namespace hrl {

constexpr int kGemmThreads = 512;          // producer / epilogue threads (16 warps)
constexpr int kGemmBlock = kGemmThreads + 32;   // + one warp that only issues tcgen05.mma
constexpr int kTileM = 128;
constexpr int kMaxN = 288;          // columns of one CTA tile (TMEM: 512 fp32 columns; shared memory: 2 stages)
constexpr int kChunkK = 32;         // reduction elements per shared-memory stage
constexpr int kStages = 3;

constexpr int kMaxSegments = 64;     // (dy, x) pairs of one segmented weight-gradient product

struct GemmOperand {
    const float *ptr, *ptr2;        // ptr2: optional second source with the same layout (operand = x*p + y*q + r), or NULL
    const float *p, *q, *r;         // per-feature constants of the operand transform, or NULL (plain operand)
    long long ld;
    int kmajor;                     // 1: element (row,k) at row*ld + k ; 0: at k*ld + row
    int relu;                       // clamp the transformed operand at 0
    int feature_is_row;             // constants indexed by the operand's row (else by the reduction index k)
    int packed;                     // B only: ptr is the hrl_board_pack image [chunk][hi|lo][n_pad rows][128 B swizzled]
};

struct GemmParams {
    GemmOperand a, b;
    const float *bias;
    float *C;
    long long ldc;
    long long c_split_stride;       // elements between the partial outputs of consecutive K slices
    int M, N, K;
    int chunks_per_split;
    int epilogue;                   // HrlGemmEpilogue
    const float *ep_y;              // masked epilogue: the pre-activation tile (M x N, leading dimension ep_ldy)
    long long ep_ldy;
    const float *ep_scale, *ep_shift, *ep_mean, *ep_rstd;     // per column, may be NULL
    float *col_partials;            // [row tiles][2][N] column sums of the epilogues that produce statistics
    int debug;                      // profiling only: 1 = no MMAs, 2 = no loads/stores
    // convolution over a board as an implicit product (no im2col in memory).  conv_off[pos * taps + tap] = (cell read by kernel
    // tap `tap` at output cell `pos`) - pos, or kConvOutside (zero padding); wrap-around boards simply have no outside.
    //   mode 1 (forward / input gradient): A rows are pixels of a channels-last tensor (ld = pixel stride), the reduction runs over
    //           (tap, channel) with every tap's channels padded to whole 32-element chunks -- chunk c reads tap c / cpt.
    //   mode 2 (weight gradient): the reduction runs over pixels, B row n is (tap, channel) = (n / cin, n % cin) of the shifted input.
    const short *conv_off;
    int conv_mode, conv_hw, conv_taps, conv_cin, conv_cpt;
    // mode 2 over several (dy, x) pairs that share one weight (a recurrent cell applied at every time step): K slice `split`
    // reads pair split / seg_splits -- ONE product per weight and backward pass instead of one per application
    int seg_splits;                 // 0 = one pair (a.ptr / b.ptr)
    int conv_ones;                  // 1: one more B row, all ones: its output column is sum_pixels dy = the bias gradient
    const float *seg_a[kMaxSegments], *seg_b[kMaxSegments];
};

constexpr short kConvOutside = -32768;
constexpr int kConvMaxTable = 256 * 9;     // cells x taps

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

// ---- B operand loaders.  Every thread owns a fixed set of "items" (one row x 4 consecutive reduction elements = one 16-byte
// shared-memory slot per split half, K-major SWIZZLE_128B: a row's chunk = 128 bytes, slot j of row r at j ^ (r & 7)); their
// coordinates are computed once, each chunk only advances the pointers.
// (An operand stored [K][rows] is transposed by the loads -- 4 scalar loads per item, coalesced along the rows.  Staging it
//  untransposed in the MN-major layout 32-bit operands have (SWIZZLE_128B_BASE32B, layout type 1, Swizzle<2,5,2> over 32 rows
//  x 4 reduction elements, LBO 4096 / SBO 512) was tried: one vector load per item and correct results, but tcgen05.mma
//  reads such an operand at 32-bit granularity and the product got 35% SLOWER, 42 -> 57 us at 288 x 288 x 16384 in 48 slices.)
struct Item {              // three registers per item (a thread holds up to 5)
    int off;               // first of the 4 elements in chunk 0, relative to the tile's first element
    uint32_t slot;         // byte offset of the 16-byte slot inside an operand half: row * 128 + ((j ^ (row & 7)) << 4)
    uint32_t meta;         // k | live << 9 | row << 10;  k = 4 * j: offset of the quad inside a chunk, row: for per-row constants
    __device__ __forceinline__ int k() const { return (int)(meta & 63u); }
    __device__ __forceinline__ bool live() const { return (meta >> 9) & 1u; }
    __device__ __forceinline__ int row() const { return (int)(meta >> 10); }
};

template <bool KMAJOR>
__device__ __forceinline__ Item make_item(int i, int n_items, long long ld, int rows_pad, int rows, int row0) {
    Item it;
    int row, j;
    if (KMAJOR) {           // 8 consecutive lanes = the 128 contiguous bytes of one row's chunk: one cache line per quarter
        row = i >> 3;       // warp in global memory, and (swizzle) 8 distinct 16-byte slots in shared memory
        j = i & 7;
    } else {                // consecutive lanes = consecutive rows: coalesced along the contiguous dimension, and the swizzle
        j = i / rows_pad;   // spreads 8 consecutive rows of one slot column over 8 distinct slots
        row = i - j * rows_pad;
    }
    const bool live = i < n_items && row < rows;
    it.meta = (uint32_t)(4 * j) | ((live ? 1u : 0u) << 9) | ((uint32_t)(row0 + row) << 10);
    it.slot = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
    it.off = (int)(KMAJOR ? (long long)row * ld + 4 * j : (long long)(4 * j) * ld + row);
    if (i >= n_items) it.slot = 0xFFFFFFFFu;
    return it;
}

template <bool KMAJOR>
__device__ __forceinline__ float4 load_item(const Item &it, const float *base, long long ld, bool vec, long long advance, int k_left) {
    // k_left = reduction elements from this chunk's start to the end of the operand (>= 32 in every chunk but the last)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!it.live()) return v;
    const float *q = base + it.off + advance;
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
    const int k = it.k();
    if (k + 0 < k_left) v.x = __ldg(q);
    if (k + 1 < k_left) v.y = __ldg(q + st);
    if (k + 2 < k_left) v.z = __ldg(q + 2 * st);
    if (k + 3 < k_left) v.w = __ldg(q + 3 * st);
    return v;
}

// weight gradient of a convolution (conv_mode 2): the item's row is (tap, channel) = (off >> 16, off & 0xFFFF), its 4 reduction
// elements are 4 consecutive pixels; each reads the pixel's tap neighbour (or nothing outside the board).  `src` is the chunk's
// table [32 pixels][taps] of source pixels (-1 = outside / past the end), computed once per chunk by the producers together.
__device__ __forceinline__ float4 load_item_conv(const Item &it, const float *base, long long ld, int k0, const GemmParams &p,
                                                 const int *src) {
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (it.live()) {
        const int ci = it.off & 0xFFFF, tap = it.off >> 16;
        if (tap >= p.conv_taps) {                    // the ones row (bias gradient)
#pragma unroll
            for (int e = 0; e < 4; e++) r[e] = (k0 + it.k() + e < p.K) ? 1.f : 0.f;
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int sp = src[(it.k() + e) * p.conv_taps + tap];
                if (sp >= 0) r[e] = __ldg(base + (long long)sp * ld + ci);
            }
        }
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}

// operand transform v = x*p[f] + y*q[f] + r[f] (relu optional) on the 4 elements of an item; elements outside the operand
// (dead rows, reduction tail) stay exactly zero.  f = the operand row, or the reduction index k0 + k + e.
__device__ __forceinline__ float4 transform_item(const GemmOperand &op, const Item &it, float4 x, float4 y, int k0, int k_left) {
    if (op.p == nullptr || !it.live()) return x;
    float4 pp, qq = make_float4(0.f, 0.f, 0.f, 0.f), rr;
    const int k = it.k();
    if (op.feature_is_row) {
        const int row = it.row();
        const float a = __ldg(op.p + row), c = __ldg(op.r + row);
        pp = make_float4(a, a, a, a);
        rr = make_float4(c, c, c, c);
        if (op.q != nullptr) { const float bq = __ldg(op.q + row); qq = make_float4(bq, bq, bq, bq); }
    } else {
        const int f = k0 + k;
        if (k + 3 < k_left) {                      // k0 and k are multiples of 4: aligned vector loads of the constants
            pp = __ldg(reinterpret_cast<const float4 *>(op.p + f));
            rr = __ldg(reinterpret_cast<const float4 *>(op.r + f));
            if (op.q != nullptr) qq = __ldg(reinterpret_cast<const float4 *>(op.q + f));
        } else {
            pp = rr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k + 0 < k_left) { pp.x = __ldg(op.p + f); rr.x = __ldg(op.r + f); if (op.q) qq.x = __ldg(op.q + f); }
            if (k + 1 < k_left) { pp.y = __ldg(op.p + f + 1); rr.y = __ldg(op.r + f + 1); if (op.q) qq.y = __ldg(op.q + f + 1); }
            if (k + 2 < k_left) { pp.z = __ldg(op.p + f + 2); rr.z = __ldg(op.r + f + 2); if (op.q) qq.z = __ldg(op.q + f + 2); }
        }
    }
    float4 v;
    v.x = fmaf(x.x, pp.x, fmaf(y.x, qq.x, rr.x));
    v.y = fmaf(x.y, pp.y, fmaf(y.y, qq.y, rr.y));
    v.z = fmaf(x.z, pp.z, fmaf(y.z, qq.z, rr.z));
    v.w = fmaf(x.w, pp.w, fmaf(y.w, qq.w, rr.w));
    if (op.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (k + 0 >= k_left) v.x = 0.f;
    if (k + 1 >= k_left) v.y = 0.f;
    if (k + 2 >= k_left) v.z = 0.f;
    if (k + 3 >= k_left) v.w = 0.f;
    return v;
}

// tcgen05.mma with the A operand in tensor memory (lane = row, one 32-bit column per reduction element) and B in shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

constexpr int kAColBase = 288;      // TMEM columns: accumulator [0, 288), A stages [288 + 64 s, ...): 32 hi + 32 lo columns each

// The A operand goes global -> registers -> (transform, hi/lo split) -> TENSOR MEMORY, the B operand -> shared memory
// (SWIZZLE_128B).  Three 3xTF32 products per k-step then read only B from shared memory: with both operands in shared
// memory the product was bound by shared-memory bandwidth (each tcgen05.mma re-read 4 KB of A and 4.6 KB of B every
// 72 cycles while the producers wrote the next stage), see profiles/README.md.
// A thread owns ONE row of the A tile (its TMEM lane: warp w may only touch lanes 32 (w % 4) ... + 31) and 8 of the 32
// reduction elements of a chunk (warp group w / 4).
// PACKED_MODE 0: B through registers; 1: B = packed image (bulk copies); 2: ... and A staged through shared memory by cp.async
// (needs 16-byte aligned A rows; two stages instead of three: the raw A tiles take the room of the third)
template <bool A_K, bool B_K, int ITEMS_A, int ITEMS_B, int PACKED_MODE>
__global__ void __launch_bounds__(kGemmBlock, 1) gemm_tf32x3_kernel(const GemmParams p, const int n_pad) {
    constexpr bool PACKED = PACKED_MODE != 0, STAGED_A = PACKED_MODE == 2;
    constexpr int NS = STAGED_A ? 2 : kStages;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // swizzle atoms need 1024-byte alignment
    __shared__ __align__(8) uint64_t bars[2 * kStages + 1];      // full[kStages] | empty[kStages] | accumulator done
    __shared__ uint32_t tmem_base_slot;
    __shared__ float b_consts[2][kMaxN];     // per-row constants of a single-source B transform (no registers, no per-chunk loads)
    __shared__ short conv_off_s[kConvMaxTable];
    __shared__ int conv_src_s[2][kChunkK * 9];      // conv_mode 2: source pixel of (pixel of the chunk, tap), two chunks in flight

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool issuer = warp == kGemmThreads / 32;
    const int m0 = blockIdx.x * kTileM;
    const int n0 = blockIdx.y * kMaxN;
    const int split = blockIdx.z;
    const int n_here = min(kMaxN, p.N - n0);
    const int rows_a = min(kTileM, p.M - m0);
    const int total_chunks = (p.K + kChunkK - 1) / kChunkK;
    const int seg = p.seg_splits ? split / p.seg_splits : 0;
    const int c_begin = (p.seg_splits ? split - seg * p.seg_splits : split) * p.chunks_per_split;
    const float *a_base = p.seg_splits ? p.seg_a[seg] : p.a.ptr, *b_base = p.seg_splits ? p.seg_b[seg] : p.b.ptr;
    const int c_end = min(total_chunks, c_begin + p.chunks_per_split);

    // shared-memory stage: [B_hi | B_lo], each [rows][128 B] with the 16-byte slots of a row swizzled
    const uint32_t b_bytes = (uint32_t)n_pad * kChunkK * 4;
    const uint32_t stage_bytes = 2 * b_bytes;
    const uint32_t smem_base = smem_u32(smem);

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(smem_u32(&bars[s]), kGemmThreads + (PACKED ? 1 : 0));   // full: every producer thread (+ the bulk copy's expect_tx)
            mbar_init(smem_u32(&bars[kStages + s]), 1);              // empty: one tcgen05.commit
        }
        mbar_init(smem_u32(&bars[2 * kStages]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.conv_mode != 0)
        for (int i = tid; i < p.conv_hw * p.conv_taps; i += kGemmBlock) conv_off_s[i] = p.conv_off[i];
    // single-source transform with per-row constants (the weight gradient's activation operand)
    const bool b_rows = !PACKED && p.b.p != nullptr && p.b.feature_is_row && p.b.ptr2 == nullptr;
    if (b_rows) {
        for (int i = tid; i < kMaxN; i += kGemmBlock) {
            const bool in = i < min(kMaxN, p.N - blockIdx.y * kMaxN);
            b_consts[0][i] = in ? __ldg(p.b.p + blockIdx.y * kMaxN + i) : 1.f;
            b_consts[1][i] = in ? __ldg(p.b.r + blockIdx.y * kMaxN + i) : 0.f;
        }
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

    // ---- A: this thread's row and 8-element slice of every chunk
    const int a_q = warp & 3, a_g = warp >> 2;
    const int a_row = a_q * 32 + lane;                   // row of the tile == TMEM lane
    const bool a_live = a_row < rows_a;
    const int a_k = 8 * a_g;                             // offset of the slice inside a chunk
    const float *a_ptr = a_base + (A_K ? (long long)(m0 + a_row) * p.a.ld + a_k : (long long)a_k * p.a.ld + (m0 + a_row));
    const long long a2 = p.a.ptr2 ? (p.a.ptr2 - p.a.ptr) : 0;
    const bool vec_a = A_K && (p.a.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a.ptr) & 15) == 0) &&
                       (!p.a.ptr2 || (reinterpret_cast<uintptr_t>(p.a.ptr2) & 15) == 0);
    float a_pr = 1.f, a_qr = 0.f, a_rr = 0.f;            // per-row transform constants
    if (p.a.p != nullptr && p.a.feature_is_row && a_live) {
        a_pr = __ldg(p.a.p + m0 + a_row);
        a_rr = __ldg(p.a.r + m0 + a_row);
        if (p.a.q != nullptr) a_qr = __ldg(p.a.q + m0 + a_row);
    }
    const uint32_t a_taddr = tmem_base + ((uint32_t)(a_q * 32) << 16) + kAColBase + a_k;

    // ---- B: items as before
    const long long off_b = B_K ? (long long)n0 * p.b.ld : (long long)n0;
    const float *Bg = b_base + off_b;
    const long long b2 = p.b.ptr2 ? (p.b.ptr2 - p.b.ptr) : 0;
    const bool vec_b = B_K && (p.b.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.b.ptr) & 15) == 0) &&
                       (!p.b.ptr2 || (reinterpret_cast<uintptr_t>(p.b.ptr2) & 15) == 0);
    const int halves = n_pad > 256 ? 2 : 1;
    const int n_mma = n_pad / halves;
    const uint32_t idesc = umma_idesc_tf32(kTileM, n_mma);
    Item ib[ITEMS_B];
#pragma unroll
    for (int u = 0; u < ITEMS_B; u++) {
        ib[u] = make_item<B_K>(tid + u * kGemmThreads, n_pad * 8, p.b.ld, n_pad, n_here, n0);
        if (!PACKED && !B_K && p.conv_mode == 2) {          // row n -> (tap, channel)
            const int n = ib[u].row(), tap = n / p.conv_cin;
            ib[u].off = (n - tap * p.conv_cin) | (tap << 16);
        }
    }

    if (issuer) {
        // ---- the MMA warp: waits for a stage to be full, issues its 3 x 4 (x halves) products, commits them to the
        //      stage's "empty" barrier (and the last ones to the accumulator barrier).  It never touches operand data.
        for (int c = c_begin; c < c_end; c++) {
            const int it = c - c_begin, s = it % NS;
            mbar_wait(smem_u32(&bars[s]), (it / NS) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                if ((p.debug & 3) != 1) {
                    const uint32_t st = smem_base + s * stage_bytes;
                    const uint32_t b_hi = st, b_lo = st + b_bytes;
                    const uint32_t a_hi = tmem_base + kAColBase + 64 * s, a_lo = a_hi + 32;
                    // (how the 288 columns are cut into instructions does not matter -- 144+144, 256+32, 192+96 all take 100 ns per
                    //  k-step and product, three instructions 127 ns: ~42 ns issue floor per instruction, scripts/gemm_mma_shapes.py)
#pragma unroll
                    for (int ks = 0; ks < kChunkK / 8; ks++) {
                        for (int h = 0; h < halves; h++) {
                            const uint32_t boff = 32 * ks + h * n_mma * 128;      // n_mma % 8 == 0: whole swizzle atoms
                            const uint32_t d = tmem_base + h * n_mma;
                            const uint32_t first = (it == 0 && ks == 0) ? 0u : 1u;
                            // small terms first: a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
                            umma_tf32_ts(d, a_lo + 8 * ks, umma_desc(b_hi + boff), idesc, first);
                            umma_tf32_ts(d, a_hi + 8 * ks, umma_desc(b_lo + boff), idesc, 1u);
                            umma_tf32_ts(d, a_hi + 8 * ks, umma_desc(b_hi + boff), idesc, 1u);
                        }
                    }
                }
                umma_commit(smem_u32(&bars[kStages + s]));
                if (c == c_end - 1) umma_commit(smem_u32(&bars[2 * kStages]));
            }
            __syncwarp();
        }
    } else {
        // ---- producers.  A: the loads of chunk c+1 are issued before chunk c is transformed / split / stored (two register
        //      sets), so a thread always has a chunk of global loads in flight.  B: staged through registers, or -- packed
        //      operand (weights pre-split and pre-swizzled by hrl_board_pack) -- ONE bulk copy per stage issued by thread 0.
        float xa[8], ya[8], xn[8], yn[8];
        auto load_a = [&](int c, float *x, float *y) {
            const int k0 = c * kChunkK, k_left = p.K - k0;
            const long long adv_a = A_K ? (long long)k0 : (long long)k0 * p.a.ld;
#pragma unroll
            for (int e = 0; e < 8; e++) x[e] = y[e] = 0.f;
            if (!a_live || (p.debug & 3) == 2) return;
            const float *q = a_ptr + adv_a;
            if (A_K && vec_a && a_k + 7 < k_left) {
                const float4 u0 = __ldg(reinterpret_cast<const float4 *>(q)), u1 = __ldg(reinterpret_cast<const float4 *>(q) + 1);
                x[0] = u0.x; x[1] = u0.y; x[2] = u0.z; x[3] = u0.w; x[4] = u1.x; x[5] = u1.y; x[6] = u1.z; x[7] = u1.w;
                if (p.a.ptr2) {
                    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(q + a2)), w1 = __ldg(reinterpret_cast<const float4 *>(q + a2) + 1);
                    y[0] = w0.x; y[1] = w0.y; y[2] = w0.z; y[3] = w0.w; y[4] = w1.x; y[5] = w1.y; y[6] = w1.z; y[7] = w1.w;
                }
            } else {
                const long long st = A_K ? 1 : p.a.ld;
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if (a_k + e < k_left) {
                        x[e] = __ldg(q + e * st);
                        if (p.a.ptr2) y[e] = __ldg(q + a2 + e * st);
                    }
            }
        };
        // STAGED_A: the chunk's A rows come through shared memory.  (Read straight from global memory a warp -- 32 rows,
        // the lanes of tensor memory -- touches 32 cache lines per load instruction, 8x the wavefronts of a coalesced copy:
        // that was what bound the producers, 0.55 us per chunk and source.)  cp.async copies them coalesced (8 lanes = one
        // row's 128 bytes) one chunk ahead into [2 stages][2 sources][128 rows][144 bytes]; a thread then reads its row's
        // 32 bytes (rows 144 bytes apart: 4 wavefronts per 512-byte request, the minimum).
        constexpr int kRawLd = 36;
        float *rawA = reinterpret_cast<float *>(smem + NS * stage_bytes);
        // (convolution: the row is a pixel, the chunk belongs to one kernel tap -> the source is the tap's neighbour pixel)
        int conv_pos[kTileM * 8 / kGemmThreads];
#pragma unroll
        for (int u = 0; u < kTileM * 8 / kGemmThreads; u++)
            conv_pos[u] = (STAGED_A && p.conv_mode == 1) ? (int)((long long)(m0 + ((tid + u * kGemmThreads) >> 3)) % p.conv_hw) : 0;
        auto issue_a = [&](int c) {
            const int st = (c - c_begin) & 1;
            const bool conv = p.conv_mode == 1;
            const int tap = conv ? c / p.conv_cpt : 0;
            const int k0 = conv ? (c - tap * p.conv_cpt) * kChunkK : c * kChunkK;      // first channel / reduction index of the chunk
            const int k_left = (conv ? p.conv_cin : p.K) - k0;
#pragma unroll
            for (int u = 0; u < kTileM * 8 / kGemmThreads; u++) {
                const int idx = tid + u * kGemmThreads, row = idx >> 3, j = idx & 7;
                int bytes = row < rows_a ? (k_left - 4 * j) * 4 : 0;
                bytes = max(0, min(16, bytes));
                long long src_row = m0 + row;
                if (conv) {
                    const short o = conv_off_s[conv_pos[u] * p.conv_taps + tap];
                    if (o == kConvOutside) bytes = 0;
                    src_row += o;
                }
                const long long off = bytes > 0 ? src_row * p.a.ld + k0 + 4 * j : 0;
                const uint32_t dst = smem_u32(rawA + ((st * 2 + 0) * kTileM + row) * kRawLd + 4 * j);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(p.a.ptr + off), "r"(bytes) : "memory");
                if (p.a.ptr2)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + kTileM * kRawLd * 4), "l"(p.a.ptr2 + off), "r"(bytes)
                                 : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (STAGED_A) {
            if (c_begin < c_end) issue_a(c_begin);
        } else if (PACKED && c_begin < c_end) {
            load_a(c_begin, xa, ya);
        }
        for (int c = c_begin; c < c_end; c++) {
            const int it = c - c_begin, s = it % NS;
            const int k0 = c * kChunkK;
            const int k_left = p.K - k0;
            const long long adv_b = B_K ? (long long)k0 : (long long)k0 * p.b.ld;
            float4 vb[ITEMS_B];
            if (!PACKED && !B_K && p.conv_mode == 2) {
                // source pixels of this chunk, once for all items: conv_src_s[it & 1][pixel][tap] (double-buffered: the readers of
                // the previous chunk are past their loads before anybody reaches this chunk's barrier)
                if (tid < kChunkK * p.conv_taps) {
                    const int kkl = tid / p.conv_taps, tap = tid - kkl * p.conv_taps, kk = k0 + kkl;
                    const short o = conv_off_s[(kk % p.conv_hw) * p.conv_taps + tap];
                    conv_src_s[it & 1][tid] = (kk < p.K && o != kConvOutside) ? kk + o : -1;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kGemmThreads) : "memory");
            }
            if (STAGED_A) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");               // my copies of this chunk have landed ...
                asm volatile("bar.sync 1, %0;" ::"n"(kGemmThreads) : "memory");    // ... everybody's; and the other raw stage is free
                if (c + 1 < c_end) issue_a(c + 1);
                const float *ra = rawA + (((it & 1) * 2 + 0) * kTileM + a_row) * kRawLd + a_k;
                const float4 u0 = reinterpret_cast<const float4 *>(ra)[0], u1 = reinterpret_cast<const float4 *>(ra)[1];
                xa[0] = u0.x; xa[1] = u0.y; xa[2] = u0.z; xa[3] = u0.w; xa[4] = u1.x; xa[5] = u1.y; xa[6] = u1.z; xa[7] = u1.w;
                if (p.a.ptr2) {
                    const float4 w0 = reinterpret_cast<const float4 *>(ra + kTileM * kRawLd)[0],
                                 w1 = reinterpret_cast<const float4 *>(ra + kTileM * kRawLd)[1];
                    ya[0] = w0.x; ya[1] = w0.y; ya[2] = w0.z; ya[3] = w0.w; ya[4] = w1.x; ya[5] = w1.y; ya[6] = w1.z; ya[7] = w1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) ya[e] = 0.f;
                }
            } else if (PACKED) {
                if (c + 1 < c_end) load_a(c + 1, xn, yn);        // next chunk's A in flight while this one is processed
            } else {
                load_a(c, xa, ya);                               // (B staged through registers: no room for a second A set)
            }
            if ((p.debug & 3) != 2 && !PACKED) {
                // all the loads first (one exposed latency per chunk, not one per item), then the transforms
#pragma unroll
                for (int u = 0; u < ITEMS_B; u++)
                    vb[u] = (!B_K && p.conv_mode == 2) ? load_item_conv(ib[u], b_base, p.b.ld, k0, p, conv_src_s[it & 1])
                                                       : load_item<B_K>(ib[u], Bg, p.b.ld, vec_b, adv_b, k_left);
                if (b_rows) {
#pragma unroll
                    for (int u = 0; u < ITEMS_B; u++) {
                        if (!ib[u].live()) continue;
                        const float pc = b_consts[0][ib[u].row() - n0], rc = b_consts[1][ib[u].row() - n0];
                        float4 v;
                        v.x = fmaf(vb[u].x, pc, rc); v.y = fmaf(vb[u].y, pc, rc); v.z = fmaf(vb[u].z, pc, rc); v.w = fmaf(vb[u].w, pc, rc);
                        if (p.b.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        if (k_left < kChunkK) {                  // the reduction tail stays exactly zero
                            const int k = ib[u].k();
                            if (k + 0 >= k_left) v.x = 0.f;
                            if (k + 1 >= k_left) v.y = 0.f;
                            if (k + 2 >= k_left) v.z = 0.f;
                            if (k + 3 >= k_left) v.w = 0.f;
                        }
                        vb[u] = v;
                    }
                } else if (p.b.p != nullptr) {
#pragma unroll
                    for (int u = 0; u < ITEMS_B; u++) {
                        const float4 y = p.b.ptr2 ? load_item<B_K>(ib[u], Bg, p.b.ld, vec_b, adv_b + b2, k_left) : make_float4(0.f, 0.f, 0.f, 0.f);
                        vb[u] = transform_item(p.b, ib[u], vb[u], y, k0, k_left);
                    }
                }
            }
            // operand transform of the current A slice
            if (p.a.p != nullptr && a_live && (p.debug & 3) != 2) {
                float pp[8], qq[8], rr[8];
                if (p.a.feature_is_row) {
#pragma unroll
                    for (int e = 0; e < 8; e++) { pp[e] = a_pr; qq[e] = a_qr; rr[e] = a_rr; }
                } else if (a_k + 7 < k_left) {      // constants of this warp's 8 reduction indices: uniform over the warp, vector loads
                    const int f = k0 + a_k;
                    const float4 p0 = __ldg(reinterpret_cast<const float4 *>(p.a.p + f)), p1 = __ldg(reinterpret_cast<const float4 *>(p.a.p + f) + 1);
                    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(p.a.r + f)), r1 = __ldg(reinterpret_cast<const float4 *>(p.a.r + f) + 1);
                    pp[0] = p0.x; pp[1] = p0.y; pp[2] = p0.z; pp[3] = p0.w; pp[4] = p1.x; pp[5] = p1.y; pp[6] = p1.z; pp[7] = p1.w;
                    rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w; rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
                    if (p.a.q) {
                        const float4 q0 = __ldg(reinterpret_cast<const float4 *>(p.a.q + f)), q1 = __ldg(reinterpret_cast<const float4 *>(p.a.q + f) + 1);
                        qq[0] = q0.x; qq[1] = q0.y; qq[2] = q0.z; qq[3] = q0.w; qq[4] = q1.x; qq[5] = q1.y; qq[6] = q1.z; qq[7] = q1.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; e++) qq[e] = 0.f;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const bool in = a_k + e < k_left;
                        pp[e] = in ? __ldg(p.a.p + k0 + a_k + e) : 0.f;
                        rr[e] = in ? __ldg(p.a.r + k0 + a_k + e) : 0.f;
                        qq[e] = (in && p.a.q) ? __ldg(p.a.q + k0 + a_k + e) : 0.f;
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    if (a_k + e < k_left) {
                        const float v = fmaf(xa[e], pp[e], fmaf(ya[e], qq[e], rr[e]));
                        xa[e] = p.a.relu ? fmaxf(v, 0.f) : v;
                    }
                }
            }
            if (it >= NS) mbar_wait(smem_u32(&bars[kStages + s]), ((it / NS) - 1) & 1);      // the MMAs that read this stage are done
            if (PACKED && tid == 0) {           // weights: one bulk copy of the stage's pre-split, pre-swizzled image
                const uint32_t bar = smem_u32(&bars[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_base + s * stage_bytes),
                             "l"(reinterpret_cast<const uint8_t *>(p.b.ptr) + (size_t)c * stage_bytes), "r"(stage_bytes), "r"(bar)
                             : "memory");
            }
            if ((p.debug & 3) != 2) {
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    hi[e] = __uint_as_float(__float_as_uint(xa[e]) & 0xFFFFE000u);
                    lo[e] = xa[e] - hi[e];
                }
                tmem_st8(a_taddr + 64 * s, hi);
                tmem_st8(a_taddr + 64 * s + 32, lo);
                if (!PACKED) {
                    uint8_t *stp = smem + s * stage_bytes;
#pragma unroll
                    for (int u = 0; u < ITEMS_B; u++) {
                        if (ib[u].slot != 0xFFFFFFFFu) {
                            float4 h4, l4;
                            split_tf32(vb[u], h4, l4);
                            *reinterpret_cast<float4 *>(stp + ib[u].slot) = h4;
                            *reinterpret_cast<float4 *>(stp + b_bytes + ib[u].slot) = l4;
                        }
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[s])) : "memory");      // this stage is full
            if (PACKED && !STAGED_A) {
#pragma unroll
                for (int e = 0; e < 8; e++) { xa[e] = xn[e]; ya[e] = yn[e]; }
            }
        }
    }

    // ---- epilogue: TMEM -> registers -> shared-memory tile (padded rows) -> coalesced global stores.
    // (a thread holds ONE row of the accumulator: storing from registers would touch 32 cache lines per warp instruction)
    //   HRL_GEMM_EP_RELU        C = max(acc, 0)
    //   HRL_GEMM_EP_STATS       C = acc, plus per-column sum and sum of squares over the tile's rows
    //   HRL_GEMM_EP_MASK_STATS  C = acc * (z > 0) with z = y*scale+shift of the pre-activation tile y (the ReLU
    //                           backward), plus per-column sums of C and of C * xhat, xhat = (y - mean) * rstd
    //                           (the two batch sums the BatchNorm backward needs)
    const int n_chunks_here = c_end - c_begin;
    float *Cg = p.C + (long long)split * p.c_split_stride;
    const int q = warp & 3, group = warp >> 2;
    constexpr int kGroups = kGemmThreads / 128;
    const int cols_per_group = ((n_pad + kGroups - 1) / kGroups + 7) / 8 * 8;
    const int ldt = n_pad + 4;                           // row stride = 16 (mod 128) bytes: conflict-free 16-byte stores
    float *tile = reinterpret_cast<float *>(smem);       // the stages are free once the last MMAs have completed
    const int ep = p.epilogue;
    // copy-out mapping: a thread owns ONE group of 4 columns (its constants and column sums live in 16 registers) and the
    // rows my_r, my_r + rpp, ...; consecutive threads = consecutive 16 bytes of a row, then of the next row
    const bool vec_c = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cg) & 15) == 0) && (n0 % 4 == 0) && (n_here % 4 == 0);
    const int cols4 = n_here >> 2;
    // rows per pass; capped by the column-sum scratch the host sized for the widest tile (a narrower last tile would take more)
    const int rpp = vec_c ? min(kGemmThreads / cols4, kGemmThreads / max(1, (n_pad - 12) / 4)) : 1;
    const int my_r = vec_c ? tid / cols4 : 0, my_c4 = tid - my_r * cols4;
    const bool mine = vec_c && !issuer && my_r < rpp;
    const bool masked = ep == HRL_GEMM_EP_MASK_STATS;
    const bool vec_y = masked && (p.ep_ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.ep_y) & 15) == 0);
    constexpr int kAhead = 4;                              // rows of the pre-activation tile in flight per thread
    float4 yq[kAhead];
    auto load_y = [&](int r) -> float4 {
        const float *yp = p.ep_y + (long long)(m0 + r) * p.ep_ldy + n0 + 4 * my_c4;
        if (vec_y) return __ldg(reinterpret_cast<const float4 *>(yp));
        return make_float4(__ldg(yp), __ldg(yp + 1), __ldg(yp + 2), __ldg(yp + 3));
    };
    if (masked && mine) {            // issued before the wait for the last MMAs: in flight while the accumulator drains
#pragma unroll
        for (int u = 0; u < kAhead; u++) {
            const int r = my_r + u * rpp;
            yq[u] = r < rows_a ? load_y(r) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (n_chunks_here > 0) {
        mbar_wait(smem_u32(&bars[2 * kStages]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (!issuer) {
        float *trow = tile + (q * 32 + (tid & 31)) * ldt;
        const uint32_t trow_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        auto put = [&](int col, const uint32_t *r, int n) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                if (e >= n) break;
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    v[i] = n_chunks_here > 0 ? __uint_as_float(r[e + i]) : 0.f;
                    if (p.bias != nullptr && col + e + i < n_here) v[i] += __ldg(p.bias + n0 + col + e + i);
                }
                *reinterpret_cast<float4 *>(trow + col + e) = make_float4(v[0], v[1], v[2], v[3]);
            }
        };
        const int col_end = min(n_pad, (group + 1) * cols_per_group);
        int col = group * cols_per_group;
        for (; col + 32 <= col_end; col += 32) {            // 32 columns per tensor-memory load (one wait per 128 bytes of a row)
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
                "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(trow_addr + (uint32_t)col));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            put(col, r, 32);
        }
        for (; col < col_end; col += 8) {
            uint32_t r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(trow_addr + (uint32_t)col));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            put(col, r, 8);
        }
    }
    __syncthreads();
    {
        const bool stats = (ep == HRL_GEMM_EP_STATS || masked) && p.col_partials != nullptr;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        if (mine) {
            float k_sc[4] = {1.f, 1.f, 1.f, 1.f}, k_sh[4] = {0.f, 0.f, 0.f, 0.f}, k_mu[4] = {0.f, 0.f, 0.f, 0.f}, k_rs[4] = {1.f, 1.f, 1.f, 1.f};
            if (masked) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int col = n0 + 4 * my_c4 + e;
                    if (p.ep_scale) k_sc[e] = __ldg(p.ep_scale + col);
                    if (p.ep_shift) k_sh[e] = __ldg(p.ep_shift + col);
                    if (p.ep_mean) k_mu[e] = __ldg(p.ep_mean + col);
                    if (p.ep_rstd) k_rs[e] = __ldg(p.ep_rstd + col);
                }
            }
            for (int r0 = my_r; r0 < rows_a; r0 += kAhead * rpp) {
#pragma unroll
                for (int u = 0; u < kAhead; u++) {
                    const int r = r0 + u * rpp;
                    if (r >= rows_a) break;
                    float4 v = reinterpret_cast<const float4 *>(tile + r * ldt)[my_c4];
                    if (ep == HRL_GEMM_EP_RELU) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    } else if (ep == HRL_GEMM_EP_STATS) {
                        s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
                        s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]);
                        s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
                    } else if (masked) {
                        const float4 y = yq[u];
                        const int rn = r + kAhead * rpp;
                        if (rn < rows_a) yq[u] = load_y(rn);              // the row this slot serves next
                        const float yv[4] = {y.x, y.y, y.z, y.w};
                        float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float z = fmaf(yv[e], k_sc[e], k_sh[e]);
                            const float d = z > 0.f ? vv[e] : 0.f;
                            const float xh = (yv[e] - k_mu[e]) * k_rs[e];
                            vv[e] = d;
                            s1[e] += d;
                            s2[e] = fmaf(d, xh, s2[e]);
                        }
                        v = make_float4(vv[0], vv[1], vv[2], vv[3]);
                    }
                    reinterpret_cast<float4 *>(Cg + (long long)(m0 + r) * p.ldc + n0)[my_c4] = v;
                }
            }
        } else if (!vec_c && !issuer) {
            for (int r = warp; r < rows_a; r += kGemmThreads / 32) {
                const float *src = tile + r * ldt;
                float *dst = Cg + (long long)(m0 + r) * p.ldc + n0;
                for (int c1 = lane; c1 < n_here; c1 += 32) dst[c1] = (ep == HRL_GEMM_EP_RELU) ? fmaxf(src[c1], 0.f) : src[c1];
            }
        }
        if (stats) {       // (the statistics epilogues require vec_c: checked by the host)
            // per-thread column sums -> shared memory (behind the tile) -> fixed-order sum over the row passes -> global partials
            float *red = tile + kTileM * ldt;                 // [rpp][2][n_pad]
            if (mine) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    red[(my_r * 2 + 0) * n_pad + 4 * my_c4 + e] = s1[e];
                    red[(my_r * 2 + 1) * n_pad + 4 * my_c4 + e] = s2[e];
                }
            }
            __syncthreads();
            for (int i = tid; i < 2 * n_here && !issuer; i += kGemmThreads) {
                const int which = i / n_here, col = i - which * n_here;
                float acc = 0.f;
                for (int w = 0; w < rpp; w++) acc += red[(w * 2 + which) * n_pad + col];
                p.col_partials[((long long)blockIdx.x * 2 + which) * p.N + n0 + col] = acc;
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

extern "C" int hrl_gemm_fused(const HrlGemmArgs *args, void *stream_) {
    using namespace hrl;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HRL_REQUIRE(args != nullptr, HRL_ERR_BAD_ARG, "hrl_gemm_fused: args is NULL");
    const HrlGemmArgs &g = *args;
    const int64_t M = g.M, N = g.N, K = g.K;
    int splits = g.splits;
    HRL_REQUIRE(g.a.ptr && g.b.ptr && (g.C || ((splits > 1 || g.segments > 0) && g.workspace)), HRL_ERR_BAD_ARG, "hrl_gemm_fused: NULL pointer");
    HRL_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 22) && K < (1ll << 31) && g.b.ld < (1ll << 22), HRL_ERR_BAD_ARG,
                "hrl_gemm_fused: bad dimensions (M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
    HRL_REQUIRE(g.conv_mode >= 0 && g.conv_mode <= 2, HRL_ERR_BAD_ARG, "hrl_gemm_fused: conv_mode is 0, 1 or 2");
    if (g.conv_mode != 0) {
        HRL_REQUIRE(g.conv_off && g.conv_hw > 0 && g.conv_taps > 0 && g.conv_cin > 0 && g.conv_cin < 65536 &&
                        (long long)g.conv_hw * g.conv_taps <= hrl::kConvMaxTable && g.a.p == nullptr && g.b.p == nullptr && g.splits >= 1,
                    HRL_ERR_BAD_ARG, "hrl_gemm_fused: convolution geometry (at most 256 cells x 9 taps, plain operands)");
        if (g.conv_mode == 1)
            HRL_REQUIRE(g.b.packed && g.a.kmajor && g.a.ld >= g.conv_cin && g.a.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(g.a.ptr) & 15) == 0 &&
                            M % g.conv_hw == 0 && K == (int64_t)g.conv_taps * ((g.conv_cin + kChunkK - 1) / kChunkK) * kChunkK && g.splits == 1,
                        HRL_ERR_UNSUPPORTED,
                        "hrl_gemm_fused: convolution forward needs a packed B image over taps x (channels padded to 32), 16-byte aligned "
                        "pixel rows and whole boards");
        else
            HRL_REQUIRE(!g.b.packed && !g.a.kmajor && !g.b.kmajor && g.b.ld >= g.conv_cin &&
                            N == (int64_t)g.conv_taps * g.conv_cin + (g.conv_ones_row ? 1 : 0) && K % g.conv_hw == 0,
                        HRL_ERR_UNSUPPORTED,
                        "hrl_gemm_fused: convolution weight gradient reduces over whole boards of pixels, N = taps x channels (+ 1 with the ones row)");
    }
    HRL_REQUIRE(g.segments >= 0 && g.segments <= hrl::kMaxSegments &&
                    (g.segments == 0 || (g.conv_mode == 2 && g.seg_a && g.seg_b && g.workspace && g.C == nullptr)),
                HRL_ERR_BAD_ARG, "hrl_gemm_fused: up to %d segments, of a convolution weight gradient left as slice partials in the workspace",
                hrl::kMaxSegments);
    HRL_REQUIRE(!g.conv_ones_row || g.conv_mode == 2, HRL_ERR_BAD_ARG, "hrl_gemm_fused: the ones row belongs to the convolution weight gradient");
    HRL_REQUIRE((g.conv_mode == 1 || g.a.ld >= (g.a.kmajor ? K : M)) && (g.b.packed || g.conv_mode == 2 || g.b.ld >= (g.b.kmajor ? K : N)) &&
                    (g.C == nullptr || g.ldc >= N),
                HRL_ERR_BAD_ARG, "hrl_gemm_fused: leading dimension smaller than the row length");
    HRL_REQUIRE(!g.a.packed && (!g.b.packed || (N <= hrl::kMaxN && g.b.p == nullptr && (reinterpret_cast<uintptr_t>(g.b.ptr) & 15) == 0)),
                HRL_ERR_BAD_ARG, "hrl_gemm_fused: only an untransformed B operand of at most 288 rows can be a packed image");
    HRL_REQUIRE((g.a.p == nullptr) == (g.a.r == nullptr) && (g.b.p == nullptr) == (g.b.r == nullptr) &&
                    (g.a.ptr2 == nullptr || (g.a.p && g.a.q)) && (g.b.ptr2 == nullptr || (g.b.p && g.b.q)),
                HRL_ERR_BAD_ARG, "hrl_gemm_fused: an operand transform needs p and r (and q with a second source)");
    HRL_REQUIRE(g.epilogue >= HRL_GEMM_EP_STORE && g.epilogue <= HRL_GEMM_EP_MASK_STATS, HRL_ERR_BAD_ARG, "hrl_gemm_fused: unknown epilogue");
    HRL_REQUIRE(g.epilogue != HRL_GEMM_EP_MASK_STATS || (g.ep_y != nullptr && g.ep_ldy >= N && (g.ep_mean == nullptr) == (g.ep_rstd == nullptr)),
                HRL_ERR_BAD_ARG, "hrl_gemm_fused: the masked epilogue needs the pre-activation tile");
    HRL_REQUIRE(g.epilogue < HRL_GEMM_EP_STATS || (N % 4 == 0 && g.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && g.col_partials),
                HRL_ERR_UNSUPPORTED, "hrl_gemm_fused: the statistics epilogues need N and ldc multiples of 4, a 16-byte aligned C and col_partials");
    const int total_chunks = (int)((K + kChunkK - 1) / kChunkK);
    if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    HRL_REQUIRE((splits == 1 && g.segments == 0) || (g.workspace != nullptr && g.bias == nullptr && g.epilogue == HRL_GEMM_EP_STORE), HRL_ERR_WORKSPACE,
                "hrl_gemm_fused: a split-K product needs a workspace of hrl_gemm_workspace_floats() floats, no bias and the plain epilogue");
    const int n_tiles = (int)((N + kMaxN - 1) / kMaxN);
    const int n_widest = (int)(N < kMaxN ? N : kMaxN);
    int n_pad = (n_widest + 15) / 16 * 16;
    if (n_pad > 256) n_pad = (n_pad + 31) / 32 * 32;

    GemmParams p;
    auto operand = [](const HrlGemmOperand &o) {
        GemmOperand r;
        r.ptr = o.ptr; r.ptr2 = o.ptr2; r.p = o.p; r.q = o.q; r.r = o.r; r.ld = o.ld;
        r.kmajor = o.kmajor ? 1 : 0; r.relu = o.relu ? 1 : 0; r.feature_is_row = o.feature_is_row ? 1 : 0; r.packed = o.packed ? 1 : 0;
        return r;
    };
    p.a = operand(g.a);
    p.b = operand(g.b);
    p.bias = g.bias;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.chunks_per_split = (total_chunks + splits - 1) / splits;
    p.epilogue = g.epilogue;
    p.ep_y = g.ep_y; p.ep_ldy = g.ep_ldy;
    p.ep_scale = g.ep_scale; p.ep_shift = g.ep_shift; p.ep_mean = g.ep_mean; p.ep_rstd = g.ep_rstd;
    p.col_partials = g.col_partials;
    p.debug = g_gemm_debug & 63;
    p.conv_off = g.conv_off; p.conv_mode = g.conv_mode; p.conv_hw = g.conv_hw; p.conv_taps = g.conv_taps; p.conv_cin = g.conv_cin;
    p.conv_cpt = (g.conv_cin + kChunkK - 1) / kChunkK;
    splits = (total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;      // no empty slices
    p.seg_splits = 0;
    p.conv_ones = g.conv_ones_row ? 1 : 0;
    if (g.segments > 0) {              // every (dy, x) pair gets `splits` K slices of its own
        p.seg_splits = splits;
        for (int i = 0; i < g.segments; i++) { p.seg_a[i] = g.seg_a[i]; p.seg_b[i] = g.seg_b[i]; }
        splits *= g.segments;
    }
    if (splits > 1 || g.segments > 0) {
        p.C = g.workspace; p.ldc = N; p.c_split_stride = M * N;
    } else {
        p.C = g.C; p.ldc = g.ldc; p.c_split_stride = 0;
    }
    const bool staged_a = (g.conv_mode == 1 || !(g_gemm_debug & 64)) && g.b.packed && g.a.kmajor && g.a.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(g.a.ptr) & 15) == 0 &&
                          (g.a.ptr2 == nullptr || (reinterpret_cast<uintptr_t>(g.a.ptr2) & 15) == 0);
    size_t smem_bytes = staged_a ? 1024 + 2 * (2 * (size_t)n_pad * kChunkK * 4) + 2 * 2 * (size_t)kTileM * 36 * 4
                                 : 1024 + (size_t)kStages * (2 * (size_t)n_pad * kChunkK * 4);
    const size_t red_rows = (size_t)kGemmThreads / (size_t)((n_pad - 12) / 4 > 0 ? (n_pad - 12) / 4 : 1);      // copy-out row passes (as the kernel)
    const size_t ep_bytes = 1024 + ((size_t)kTileM * (n_pad + 4) + 2 * red_rows * (size_t)n_pad) * 4;      // epilogue tile + column-sum scratch
    if (smem_bytes < ep_bytes) smem_bytes = ep_bytes;
    const dim3 grid((unsigned)((M + kTileM - 1) / kTileM), (unsigned)n_tiles, (unsigned)splits);
    const int items_b = (n_pad * 8 + kGemmThreads - 1) / kGemmThreads;
    constexpr int IA = kTileM * 8 / kGemmThreads;
#define HRL_GEMM_LAUNCH3(AK, BK, IB, PK)                                                                                   \
    {                                                                                                                     \
        HRL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<AK, BK, IA, IB, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            (int)smem_bytes));                                                            \
        gemm_tf32x3_kernel<AK, BK, IA, IB, PK><<<grid, kGemmBlock, smem_bytes, stream>>>(p, n_pad);                      \
    }
#define HRL_GEMM_LAUNCH2(AK, BK, IB) HRL_GEMM_LAUNCH3(AK, BK, IB, 0)
#define HRL_GEMM_LAUNCH(IB)                                                                    \
    {                                                                                         \
        if (p.a.kmajor && p.b.kmajor) HRL_GEMM_LAUNCH2(true, true, IB)                        \
        else if (p.a.kmajor) HRL_GEMM_LAUNCH2(true, false, IB)                                \
        else if (p.b.kmajor) HRL_GEMM_LAUNCH2(false, true, IB)                                \
        else HRL_GEMM_LAUNCH2(false, false, IB)                                               \
    }
    if (staged_a) HRL_GEMM_LAUNCH3(true, true, 1, 2)
    else if (p.b.packed) {
        if (p.a.kmajor) HRL_GEMM_LAUNCH3(true, true, 1, 1)
        else HRL_GEMM_LAUNCH3(false, true, 1, 1)
    } else if (items_b <= 1) HRL_GEMM_LAUNCH(1)
    else if (items_b <= 3) HRL_GEMM_LAUNCH(3)
    else HRL_GEMM_LAUNCH(5)
#undef HRL_GEMM_LAUNCH
#undef HRL_GEMM_LAUNCH2
#undef HRL_GEMM_LAUNCH3
    HRL_CUDA_CHECK(cudaGetLastError());
    if (splits > 1 && g.C != nullptr) {       // C == NULL: the caller consumes the slice partials itself (hrl_board_fold)
        const long long n = (long long)M * N;
        HRL_REQUIRE(g.ldc == N, HRL_ERR_UNSUPPORTED, "hrl_gemm_fused: split-K output must be dense (ldc == N)");
        int blocks = (int)((n + 255) / 256);
        if (blocks > 1184) blocks = 1184;
        sum_partials_kernel<<<blocks, 256, 0, stream>>>(g.workspace, splits, n, n, g.C);
        HRL_CUDA_CHECK(cudaGetLastError());
    }
    return HRL_OK;
}

extern "C" int hrl_gemm_tf32x3(const float *A, int64_t lda, int32_t a_kmajor, const float *B, int64_t ldb, int32_t b_kmajor,
                               const float *bias, float *C, int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t splits,
                               float *workspace, void *stream) {
    HrlGemmArgs g;
    memset(&g, 0, sizeof(g));
    g.a.ptr = A; g.a.ld = lda; g.a.kmajor = a_kmajor;
    g.b.ptr = B; g.b.ld = ldb; g.b.kmajor = b_kmajor;
    g.bias = bias; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.splits = splits; g.workspace = workspace;
    g.epilogue = HRL_GEMM_EP_STORE;
    return hrl_gemm_fused(&g, stream);
}
