// K2 -- replay gather/pad: the device form of make_batch (handyrl/train.py:33-124).
//
// Episodes live decoded in a flat "replay store" in HBM (one row per step, see HrlGatherArgs);
// a batch is B window descriptors.  One CTA writes `cells_per_block` consecutive (b,t) cells of
// every batch tensor: live cells copy from the store, cells outside the window get the pad
// constants of train.py:92-106 (prob 1, action_mask 1e32, progress 1, value = outcome after
// the end, everything else 0).  Every batch byte is written exactly once, coalesced along the
// innermost dimension; the store is read once.
#include "common.cuh"

namespace hrl {

struct CellSrc {
    int64_t row;   // store row of this cell's step, or -1 when the cell is padding
    int after;     // padding after the end of the window (value = outcome, train.py:98)
    int step;      // episode step index
};

__device__ __forceinline__ CellSrc locate(const HrlWindow &w, int t, int burn_in) {
    CellSrc c;
    const int len = w.end - w.start;
    const int pad_b = burn_in - (w.train_start - w.start);  // train.py:94
    const int k = t - pad_b;
    c.step = w.start + k;
    c.after = (k >= len);
    c.row = (k >= 0 && k < len) ? (w.first_step + c.step) : -1;
    return c;
}

// copy n floats src -> dst (or fill when src == nullptr) by one warp, 16 bytes per lane when both sides allow it
__device__ __forceinline__ void warp_copy_row(float *__restrict__ dst, const float *__restrict__ src, float fill, int n, int lane) {
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                     (src == nullptr || (reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
        const int n4 = n >> 2;
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        if (src) {
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
#pragma unroll 4
            for (int i = lane; i < n4; i += 32) __stcs(d4 + i, __ldcs(s4 + i));
        } else {
            const float4 f4 = make_float4(fill, fill, fill, fill);
            for (int i = lane; i < n4; i += 32) __stcs(d4 + i, f4);
        }
    } else {
        if (src) for (int i = lane; i < n; i += 32) dst[i] = src[i];
        else for (int i = lane; i < n; i += 32) dst[i] = fill;
    }
}

// One warp per (cell, policy row): it locates the source step once and streams the observation and the
// action-mask row (vectorised); lanes then write the handful of per-player scalars of the cell.
__global__ void __launch_bounds__(256) gather_pad_kernel(const HrlGatherArgs g, int cells_per_block) {
    const int T = g.T, P = g.P, Pa = g.Pa, A = g.A, Ps = g.Ps, OE = g.obs_elems;
    const int64_t ncell = (int64_t)g.B * T;
    const int64_t c0 = (int64_t)blockIdx.x * cells_per_block;
    const int nc = (int)min((int64_t)cells_per_block, ncell - c0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const bool solo = (P == 1 && Ps > 1);

    for (int u = warp; u < nc * Pa; u += nwarp) {
        const int c = u / Pa, q = u - c * Pa;
        const int64_t cell = c0 + c;
        const int b = (int)(cell / T), t = (int)(cell - (int64_t)b * T);
        const HrlWindow w = g.windows[b];
        const CellSrc s = locate(w, t, g.burn_in);
        const bool live = s.row >= 0;
        // policy-side player of row q in this cell (train.py:65-68)
        const int pl = !live ? 0 : (g.turn_alternating ? g.st_turn[s.row] : (solo ? w.player : q));
        const int64_t sp = live ? s.row * Ps + pl : 0;
        if (OE > 0) warp_copy_row(g.observation + (cell * Pa + q) * OE, live ? g.st_obs + sp * OE : nullptr, 0.0f, OE, lane);
        warp_copy_row(g.action_mask + (cell * Pa + q) * A, live ? g.st_amask + sp * A : nullptr, 1e32f, A, lane);
        if (lane == 0) {
            g.selected_prob[cell * Pa + q] = live ? g.st_prob[sp] : 1.0f;
            g.action[cell * Pa + q] = live ? (int64_t)g.st_action[sp] : 0;
        }
        if (q == 0) {   // the cell's value-side and per-cell scalars, once
            for (int p = lane; p < P; p += 32) {
                const int vp = solo ? w.player : p;
                float val = 0.f, rew = 0.f, ret = 0.f, tm = 0.f, om = 0.f;
                if (live) {
                    const int64_t sv = s.row * Ps + vp;
                    val = g.st_value[sv];
                    rew = g.st_reward[sv];
                    ret = g.st_return[sv];
                    const uint8_t f = g.st_flags[sv];
                    tm = (f & 1) ? 1.0f : 0.0f;
                    om = (f & 2) ? 1.0f : 0.0f;
                } else if (s.after) {
                    val = g.st_outcome[(int64_t)w.outcome_row * Ps + vp];  // np.tile(oc, ...) train.py:98
                }
                const int64_t o = cell * P + p;
                g.value[o] = val;
                g.reward[o] = rew;
                g.ret[o] = ret;
                g.turn_mask[o] = tm;
                g.observation_mask[o] = om;
                if (t == 0) g.outcome[(int64_t)b * P + p] = g.st_outcome[(int64_t)w.outcome_row * Ps + vp];
            }
            if (lane == 0) {
                g.episode_mask[cell] = live ? 1.0f : 0.0f;
                g.progress[cell] = live ? (float)s.step / (float)w.total : 1.0f;  // train.py:89, 106
            }
        }
    }
}

}  // namespace hrl

extern "C" int hrl_gather_pad(const HrlGatherArgs *args, void *stream) {
    using namespace hrl;
    HRL_REQUIRE(args != nullptr, HRL_ERR_BAD_ARG, "hrl_gather_pad: args is NULL");
    const HrlGatherArgs &g = *args;
    HRL_REQUIRE(g.B > 0 && g.T > 0 && g.P > 0 && g.A > 0 && g.Ps > 0 && g.obs_elems >= 0, HRL_ERR_BAD_ARG,
                "hrl_gather_pad: non-positive dimension");
    HRL_REQUIRE(g.Pa == 1 || g.Pa == g.P, HRL_ERR_BAD_ARG, "hrl_gather_pad: Pa must be 1 or P");
    HRL_REQUIRE(g.P == g.Ps || g.P == 1, HRL_ERR_BAD_ARG, "hrl_gather_pad: P must equal Ps, or 1 for solo training");
    HRL_REQUIRE(g.burn_in >= 0 && g.burn_in < g.T, HRL_ERR_BAD_ARG, "hrl_gather_pad: burn_in outside [0,T)");
    HRL_REQUIRE(g.windows && g.st_prob && g.st_action && g.st_amask && g.st_value && g.st_reward && g.st_return &&
                    g.st_flags && g.st_outcome && (g.obs_elems == 0 || g.st_obs) && (!g.turn_alternating || g.st_turn),
                HRL_ERR_BAD_ARG, "hrl_gather_pad: a replay-store pointer is NULL");
    HRL_REQUIRE(g.selected_prob && g.value && g.action && g.outcome && g.reward && g.ret && g.episode_mask &&
                    g.turn_mask && g.observation_mask && g.action_mask && g.progress && (g.obs_elems == 0 || g.observation),
                HRL_ERR_BAD_ARG, "hrl_gather_pad: a batch output pointer is NULL");
    // 8 warps per CTA, one (cell, policy row) per warp at a time; a CTA takes enough cells for ~16 KB of copies
    const int64_t per_cell = (int64_t)g.Pa * (g.obs_elems + g.A) + 2 * g.Pa + 5 * g.P + 2;
    int cpb = (int)(4096 / per_cell);
    if (cpb < (8 + g.Pa - 1) / g.Pa) cpb = (8 + g.Pa - 1) / g.Pa;
    const int64_t ncell = (int64_t)g.B * g.T;
    // keep at least ~4 CTAs per SM in flight when the batch is small
    while (cpb > 1 && (ncell + cpb - 1) / cpb < 4 * kNumSM) cpb >>= 1;
    const int64_t grid = (ncell + cpb - 1) / cpb;
    gather_pad_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, cpb);
    HRL_CUDA_CHECK(cudaGetLastError());
    return HRL_OK;
}
