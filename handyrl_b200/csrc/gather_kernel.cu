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

__global__ void __launch_bounds__(256) gather_pad_kernel(const HrlGatherArgs g, int cells_per_block) {
    const int T = g.T, P = g.P, Pa = g.Pa, A = g.A, Ps = g.Ps, OE = g.obs_elems;
    const int64_t ncell = (int64_t)g.B * T;
    const int64_t c0 = (int64_t)blockIdx.x * cells_per_block;
    const int nc = (int)min((int64_t)cells_per_block, ncell - c0);
    const int tid = threadIdx.x, nthr = blockDim.x;

    // policy-side player of row q in this cell (train.py:65-68)
#define POLICY_PLAYER(w, src, q) (g.turn_alternating ? g.st_turn[(src).row] : ((P == 1 && Ps > 1) ? (w).player : (q)))
#define VALUE_PLAYER(w, p) ((P == 1 && Ps > 1) ? (w).player : (p))

    // observation (B,T,Pa,OE)
    for (int i = tid; i < nc * Pa * OE; i += nthr) {
        const int c = i / (Pa * OE), r = i - c * (Pa * OE), q = r / OE, j = r - q * OE;
        const int64_t cell = c0 + c;
        const int b = (int)(cell / T), t = (int)(cell - (int64_t)b * T);
        const HrlWindow w = g.windows[b];
        const CellSrc s = locate(w, t, g.burn_in);
        float v = 0.0f;
        if (s.row >= 0) v = g.st_obs[(s.row * Ps + POLICY_PLAYER(w, s, q)) * OE + j];
        g.observation[cell * Pa * OE + r] = v;
    }
    // action_mask (B,T,Pa,A)
    for (int i = tid; i < nc * Pa * A; i += nthr) {
        const int c = i / (Pa * A), r = i - c * (Pa * A), q = r / A, j = r - q * A;
        const int64_t cell = c0 + c;
        const int b = (int)(cell / T), t = (int)(cell - (int64_t)b * T);
        const HrlWindow w = g.windows[b];
        const CellSrc s = locate(w, t, g.burn_in);
        float v = 1e32f;
        if (s.row >= 0) v = g.st_amask[(s.row * Ps + POLICY_PLAYER(w, s, q)) * A + j];
        g.action_mask[cell * Pa * A + r] = v;
    }
    // policy-side scalars (B,T,Pa)
    for (int i = tid; i < nc * Pa; i += nthr) {
        const int c = i / Pa, q = i - c * Pa;
        const int64_t cell = c0 + c;
        const int b = (int)(cell / T), t = (int)(cell - (int64_t)b * T);
        const HrlWindow w = g.windows[b];
        const CellSrc s = locate(w, t, g.burn_in);
        float prob = 1.0f;
        int64_t act = 0;
        if (s.row >= 0) {
            const int64_t sp = s.row * Ps + POLICY_PLAYER(w, s, q);
            prob = g.st_prob[sp];
            act = g.st_action[sp];
        }
        g.selected_prob[cell * Pa + q] = prob;
        g.action[cell * Pa + q] = act;
    }
    // value-side scalars (B,T,P)
    for (int i = tid; i < nc * P; i += nthr) {
        const int c = i / P, p = i - c * P;
        const int64_t cell = c0 + c;
        const int b = (int)(cell / T), t = (int)(cell - (int64_t)b * T);
        const HrlWindow w = g.windows[b];
        const CellSrc s = locate(w, t, g.burn_in);
        const int pl = VALUE_PLAYER(w, p);
        float val = 0.f, rew = 0.f, ret = 0.f, tm = 0.f, om = 0.f;
        if (s.row >= 0) {
            const int64_t sp = s.row * Ps + pl;
            val = g.st_value[sp];
            rew = g.st_reward[sp];
            ret = g.st_return[sp];
            const uint8_t f = g.st_flags[sp];
            tm = (f & 1) ? 1.0f : 0.0f;
            om = (f & 2) ? 1.0f : 0.0f;
        } else if (s.after) {
            val = g.st_outcome[(int64_t)w.outcome_row * Ps + pl];  // np.tile(oc, ...) train.py:98
        }
        const int64_t o = cell * P + p;
        g.value[o] = val;
        g.reward[o] = rew;
        g.ret[o] = ret;
        g.turn_mask[o] = tm;
        g.observation_mask[o] = om;
    }
    // per-cell scalars (B,T)
    for (int i = tid; i < nc; i += nthr) {
        const int64_t cell = c0 + i;
        const int b = (int)(cell / T), t = (int)(cell - (int64_t)b * T);
        const HrlWindow w = g.windows[b];
        const CellSrc s = locate(w, t, g.burn_in);
        g.episode_mask[cell] = s.row >= 0 ? 1.0f : 0.0f;
        g.progress[cell] = s.row >= 0 ? (float)s.step / (float)w.total : 1.0f;  // train.py:89, 106
        if (t == 0)
            for (int p = 0; p < P; p++)
                g.outcome[(int64_t)b * P + p] = g.st_outcome[(int64_t)w.outcome_row * Ps + VALUE_PLAYER(w, p)];
    }
#undef POLICY_PLAYER
#undef VALUE_PLAYER
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
    const int64_t per_cell = (int64_t)g.Pa * (g.obs_elems + g.A) + 2 * g.Pa + 5 * g.P + 2;
    int cpb = (int)(4096 / per_cell);
    if (cpb < 1) cpb = 1;
    const int64_t ncell = (int64_t)g.B * g.T;
    // keep at least ~4 CTAs per SM in flight when the batch is small
    while (cpb > 1 && (ncell + cpb - 1) / cpb < 4 * kNumSM) cpb >>= 1;
    const int64_t grid = (ncell + cpb - 1) / cpb;
    gather_pad_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, cpb);
    HRL_CUDA_CHECK(cudaGetLastError());
    return HRL_OK;
}
