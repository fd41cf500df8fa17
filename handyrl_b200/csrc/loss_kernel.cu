// K1 -- fused forward+backward of the HandyRL loss over one replay batch (sm_100a).
//
// Replaces, in ONE launch and one pass over HBM:
//   handyrl/train.py:176-184  mask epilogue of forward_prediction
//   handyrl/train.py:218-267  compute_loss (log-softmax/gather, importance ratio clamp,
//                             2-player value symmetrisation, targets/advantages)
//   handyrl/losses.py:16-80   MC / TD(lambda) / UPGO / V-Trace reverse-time scans
//   handyrl/train.py:189-215  compose_losses (masked sums, entropy)
//   + the autograd pass of train.py:369 through those ops (closed-form gradients
//     w.r.t. the raw policy / value / return outputs of the net).
//
// Mapping (see DESIGN.md "K1"):
//   CTA        = EPB consecutive episodes (windows) b; all of their T x P x A data stays on chip
//   phase 0    = small per-cell tensors -> shared memory, coalesced
//   phase 1    = one group of LPR lanes per policy row (b,t,pa): masked logits z = raw*scale - amask,
//                row max / sum-exp / entropy / gather with warp shuffles; z is staged in shared memory
//   phase 2    = one thread per (b,p) column runs every reverse-time recurrence over T in registers
//                (value + return stream, target + advantage algorithm), emits per-cell terms
//   phase 3    = row groups again: dL/dpolicy_raw from the staged z (no second HBM read),
//                dL/dvalue_raw, dL/dreturn_raw, coalesced streaming stores
//   phase 4    = block partials -> workspace; the last CTA to finish reduces them in a fixed
//                order in fp64 (deterministic) and writes the 6 scalars.
#include "common.cuh"
#include <math.h>

namespace hrl {

struct LossParams {
    HrlLossArgs a;
    int Tt;       // trained steps = T - burn_in
    int EPB;      // episodes per CTA
    int stage_z;  // masked logits kept in shared memory between phase 1 and 3
    int has_v, has_r;
};

// shared-memory carve-up, in floats
struct SmemLayout {
    int emask, prog;                               // [cells]
    int tm, om, rew, ret, wterm, dv, dr;           // [cols]
    int outcome;                                   // [EPB*P]
    int logp, rho, ent, mx, lsum, scale, vraw, rraw, act;  // [rows]
    int red;                                       // [8*32]
    int z;                                         // [rows*A] if staged
    int total;
};

__host__ __device__ inline SmemLayout make_layout(int EPB, int Tt, int P, int Pa, int A, int stage_z) {
    SmemLayout L;
    int cells = EPB * Tt, cols = cells * P, rows = cells * Pa, o = 0;
    L.emask = o; o += cells;
    L.prog = o; o += cells;
    L.tm = o; o += cols;
    L.om = o; o += cols;
    L.rew = o; o += cols;
    L.ret = o; o += cols;
    L.wterm = o; o += cols;
    L.dv = o; o += cols;
    L.dr = o; o += cols;
    L.outcome = o; o += EPB * P;
    L.logp = o; o += rows;
    L.rho = o; o += rows;
    L.ent = o; o += rows;
    L.mx = o; o += rows;
    L.lsum = o; o += rows;
    L.scale = o; o += rows;
    L.vraw = o; o += rows;
    L.rraw = o; o += rows;
    L.act = o; o += rows;
    o = (o + 3) & ~3;
    L.red = o; o += 8 * 32;
    L.z = o;
    if (stage_z) o += rows * A;
    L.total = o;
    return L;
}

// One reverse-time step of one recurrence (losses.py:16-60).  `st*` carry the values of step t+1.
struct Chain {
    float G;        // TD / UPGO target at t+1
    float acc;      // V-Trace vs - v at t+1
    float vs_next;  // V-Trace vs at t+1
};

__device__ __forceinline__ void chain_step(int algo, bool has_baseline, bool last, float v_t, float v_next,
                                           float lam_next, float r_t, float gamma, float boot, float ret_t,
                                           float rho_t, float c_t, Chain &s, float &tgt, float &adv) {
    if (!has_baseline) {  // losses.py:64-66
        tgt = ret_t;
        adv = ret_t;
        return;
    }
    switch (algo) {
        case HRL_MC:  // losses.py:16-17
            tgt = ret_t;
            adv = ret_t - v_t;
            break;
        case HRL_TD:  // losses.py:20-29
            s.G = last ? boot : r_t + gamma * ((1.0f - lam_next) * v_next + lam_next * s.G);
            tgt = s.G;
            adv = s.G - v_t;
            break;
        case HRL_UPGO:  // losses.py:32-42
            s.G = last ? boot : r_t + gamma * fmaxf(v_next, (1.0f - lam_next) * v_next + lam_next * s.G);
            tgt = s.G;
            adv = s.G - v_t;
            break;
        default: {  // HRL_VTRACE, losses.py:45-60
            float vn = last ? boot : v_next;
            float delta = rho_t * (r_t + gamma * vn - v_t);
            s.acc = last ? delta : delta + gamma * lam_next * c_t * s.acc;
            float vs = s.acc + v_t;
            adv = r_t + gamma * (last ? boot : s.vs_next) - v_t;
            s.vs_next = vs;
            tgt = vs;
        }
    }
}

template <int LPR, int NPL>
__global__ void __launch_bounds__(512) loss_fwd_bwd_kernel(const LossParams prm) {
    extern __shared__ __align__(16) float smem[];
    const HrlLossArgs &a = prm.a;
    const int T0 = a.T, P = a.P, Pa = a.Pa, A = a.A, bi = a.burn_in, Tt = prm.Tt;
    const int b0 = blockIdx.x * prm.EPB;
    const int nE = min(prm.EPB, a.B - b0);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const SmemLayout L = make_layout(prm.EPB, Tt, P, Pa, A, prm.stage_z);
    const int R = Tt * Pa;           // rows per episode
    const int nrows = nE * R, ncols = nE * Tt * P, ncells = nE * Tt;
    int *s_act = reinterpret_cast<int *>(smem + L.act);

    // ---------------- phase 0: small tensors -> smem (coalesced; each episode's slice is contiguous)
    for (int i = tid; i < ncols; i += nthr) {
        int e = i / (Tt * P), r = i - e * (Tt * P);
        size_t g = ((size_t)(b0 + e) * T0 + bi) * P + r;
        smem[L.tm + i] = a.turn_mask[g];
        smem[L.om + i] = a.observation_mask[g];
        smem[L.rew + i] = a.reward[g];
        smem[L.ret + i] = a.ret[g];
    }
    for (int i = tid; i < ncells; i += nthr) {
        int e = i / Tt, r = i - e * Tt;
        size_t g = (size_t)(b0 + e) * T0 + bi + r;
        smem[L.emask + i] = a.episode_mask[g];
        smem[L.prog + i] = a.progress[g];
    }
    for (int i = tid; i < nrows; i += nthr) {
        int e = i / R, r = i - e * R;
        size_t g = ((size_t)(b0 + e) * T0 + bi) * Pa + r;
        smem[L.vraw + i] = prm.has_v ? a.value_raw[g] : 0.0f;
        smem[L.rraw + i] = prm.has_r ? a.return_raw[g] : 0.0f;
        s_act[i] = (int)a.action[g];
    }
    for (int i = tid; i < nE * P; i += nthr) smem[L.outcome + i] = a.outcome[(size_t)b0 * P + i];

    // ---------------- phase 1: per-row softmax statistics
    const int grp = tid / LPR, lane = tid % LPR, ngrp = nthr / LPR;
    for (int base = 0; base < nrows; base += ngrp) {
        const int r = base + grp;
        const bool valid = r < nrows;
        float z[NPL];
        float scale = 0.0f, em = 0.0f, mu = 1.0f;
        int act = 0;
        size_t grow = 0;
        if (valid) {
            int e = r / R, rr = r - e * R, t = rr / Pa, q = rr - t * Pa;
            size_t cell = (size_t)(b0 + e) * T0 + bi + t;
            grow = cell * Pa + q;
            if (Pa == P) {
                scale = a.turn_mask[cell * P + q];
            } else {  // turn-alternating batch: sum over players (train.py:179-180)
                for (int p = 0; p < P; p++) scale += a.turn_mask[cell * P + p];
            }
            em = a.episode_mask[cell];
            mu = a.selected_prob[grow];
            act = (int)a.action[grow];
        }
        const float *rawp = a.policy_raw + grow * A;
        const float *amp = a.action_mask + grow * A;
        float m = -INFINITY, za = -INFINITY;
#pragma unroll
        for (int k = 0; k < NPL; k++) {
            int j = k * LPR + lane;
            z[k] = -INFINITY;
            if (valid && j < A) {
                z[k] = ld_stream(rawp + j) * scale - ld_stream(amp + j);  // train.py:178-181
                if (j == act) za = z[k];
            }
            m = fmaxf(m, z[k]);
        }
        m = group_max<LPR>(m);
        za = group_max<LPR>(za);
        float se = 0.0f;
#pragma unroll
        for (int k = 0; k < NPL; k++) {
            int j = k * LPR + lane;
            if (j < A) se += expf(z[k] - m);
        }
        se = group_sum<LPR>(se);
        const float lsum = logf(se);
        float h = 0.0f;
#pragma unroll
        for (int k = 0; k < NPL; k++) {
            int j = k * LPR + lane;
            if (valid && j < A) {
                float lp = z[k] - m - lsum;
                float pj = expf(lp);
                h -= pj * fmaxf(lp, -3.402823466e38f);  // Categorical.entropy clamps logits at finfo.min
                if (prm.stage_z) smem[L.z + (size_t)r * A + j] = z[k];
            }
        }
        h = group_sum<LPR>(h);
        if (valid && lane == 0) {
            float lt = (za - m - lsum) * em;                              // train.py:232
            float lb = logf(fminf(fmaxf(mu, 1e-16f), 1.0f)) * em;         // train.py:231
            float rho = fminf(fmaxf(expf(lt - lb), 0.0f), 1.0f);          // train.py:235-238
            smem[L.logp + r] = lt;
            smem[L.rho + r] = rho;
            smem[L.ent + r] = h;
            smem[L.mx + r] = m;
            smem[L.lsum + r] = lsum;
            smem[L.scale + r] = scale;
            if (a.tap_logp) a.tap_logp[grow] = lt;
            if (a.tap_rho) a.tap_rho[grow] = rho;
            if (a.tap_entropy) a.tap_entropy[grow] = h;
        }
    }
    __syncthreads();

    // ---------------- phase 2: reverse-time recurrences, one thread per (episode, player) column
    float Lp = 0.f, Lv = 0.f, Lr = 0.f, Lent = 0.f, Lreg = 0.f, dcnt = 0.f;
    if (tid < nE * P) {
        const int e = tid / P, p = tid - e * P;
        const int q = (Pa == P) ? p : 0;
        const bool sym = a.two_player_zero_sum && P == 2;
        const int po = sym ? 1 - p : p, qo = (Pa == P) ? po : 0;
        const float lmb = a.lambda, gam = a.gamma, dec = a.entropy_regularization_decay;
        const float oc = smem[L.outcome + e * P + p];
        const int vt = a.value_target, pt = a.policy_target;
        const bool two = (pt != vt);
        const size_t gcol0 = ((size_t)(b0 + e) * T0 + bi) * P + p;
        const float boot_r = smem[L.ret + (e * Tt + Tt - 1) * P + p];  // returns[:, -1]
        Chain cv = {0, 0, 0}, cv2 = {0, 0, 0}, cr = {0, 0, 0}, cr2 = {0, 0, 0};
        float v_next = 0.f, lamv_next = 0.f, r_next = 0.f, lamr_next = 0.f;
        for (int t = Tt - 1; t >= 0; t--) {
            const int cell = e * Tt + t, col = cell * P + p, row = cell * Pa + q;
            const float om = smem[L.om + col], tm = smem[L.tm + col], em = smem[L.emask + cell];
            const float vout = smem[L.vraw + row] * om;      // train.py:184
            const float rout = smem[L.rraw + row] * om;
            float vb = vout, vm = om;
            if (sym) {  // train.py:243-247
                const float omo = smem[L.om + cell * P + po];
                const float vo = -(smem[L.vraw + cell * Pa + qo] * omo);
                vb = (vout * om + vo * omo) / (om + omo + 1e-8f);
                vm = fminf(fmaxf(om + omo, 0.0f), 1.0f);
            }
            vb = vb * em + oc * (1.0f - em);                 // train.py:248
            const float lamv = lmb + (1.0f - lmb) * (1.0f - vm);   // losses.py:71
            const float lamr = lmb + (1.0f - lmb) * (1.0f - om);
            const float rho = smem[L.rho + row];
            const float rew = smem[L.rew + col], ret = smem[L.ret + col];
            const bool last = (t == Tt - 1);
            float tgv, adv_v, tgr, adv_r, dummy;
            chain_step(vt, prm.has_v, last, vb, v_next, lamv_next, 0.0f, 1.0f, oc, oc, rho, rho, cv, tgv, adv_v);
            chain_step(vt, prm.has_r, last, rout, r_next, lamr_next, rew, gam, boot_r, ret, rho, rho, cr, tgr, adv_r);
            if (two) {  // train.py:260-262
                chain_step(pt, prm.has_v, last, vb, v_next, lamv_next, 0.0f, 1.0f, oc, oc, rho, rho, cv2, dummy, adv_v);
                chain_step(pt, prm.has_r, last, rout, r_next, lamr_next, rew, gam, boot_r, ret, rho, rho, cr2, dummy, adv_r);
            }
            v_next = vb; lamv_next = lamv; r_next = rout; lamr_next = lamr;

            const float tot_adv = rho * (adv_v + adv_r);     // train.py:265
            smem[L.wterm + col] = tot_adv * tm;
            Lp += -smem[L.logp + row] * tot_adv * tm;        // train.py:202
            float dv = 0.f, dr = 0.f;
            if (prm.has_v) {                                  // train.py:204
                float d = vout - tgv;
                Lv += d * d * om;
                dv = d * om * om;
            }
            if (prm.has_r) {                                  // train.py:206 smooth_l1, beta 1
                float d = rout - tgr, ad = fabsf(d);
                Lr += (ad < 1.0f ? 0.5f * d * d : ad - 0.5f) * om;
                dr = fminf(fmaxf(d, -1.0f), 1.0f) * om * om;
            }
            smem[L.dv + col] = dv;
            smem[L.dr + col] = dr;
            const float h = smem[L.ent + row] * tm;          // train.py:208
            Lent += h;
            Lreg += h * (1.0f - smem[L.prog + cell] * (1.0f - dec));   // train.py:212
            dcnt += tm;
            const size_t gcol = gcol0 + (size_t)t * P;
            if (a.tap_target_value) a.tap_target_value[gcol] = tgv;
            if (a.tap_target_return) a.tap_target_return[gcol] = tgr;
            if (a.tap_advantage) a.tap_advantage[gcol] = tot_adv;
        }
    }
    __syncthreads();

    // ---------------- phase 3: gradients w.r.t. the raw net outputs
    const float creg = a.entropy_regularization;
    for (int base = 0; base < nrows; base += ngrp) {
        const int r = base + grp;
        if (r >= nrows) continue;  // no shuffles below: divergence is harmless
        const int e = r / R, rr = r - e * R, t = rr / Pa, q = rr - t * Pa;
        const int cell = e * Tt + t;
        const size_t grow = ((size_t)(b0 + e) * T0 + bi + t) * Pa + q;
        float w = 0.f, k = 0.f, gv = 0.f, gr = 0.f;
        if (Pa == P) {
            w = smem[L.wterm + cell * P + q];
            k = smem[L.tm + cell * P + q];
            gv = smem[L.dv + cell * P + q];
            gr = smem[L.dr + cell * P + q];
        } else {
            for (int p = 0; p < P; p++) {
                w += smem[L.wterm + cell * P + p];
                k += smem[L.tm + cell * P + p];
                gv += smem[L.dv + cell * P + p];
                gr += smem[L.dr + cell * P + p];
            }
        }
        w *= smem[L.emask + cell];
        k *= creg * (1.0f - smem[L.prog + cell] * (1.0f - a.entropy_regularization_decay));
        const float scale = smem[L.scale + r], m = smem[L.mx + r], lsum = smem[L.lsum + r], h = smem[L.ent + r];
        const int act = s_act[r];
        float *outp = a.dpolicy_raw + grow * A;
        if (scale == 0.0f) {
#pragma unroll
            for (int kk = 0; kk < NPL; kk++) {
                int j = kk * LPR + lane;
                if (j < A) st_stream(outp + j, 0.0f);
            }
        } else {
            const float *rawp = a.policy_raw + grow * A;
            const float *amp = a.action_mask + grow * A;
#pragma unroll
            for (int kk = 0; kk < NPL; kk++) {
                int j = kk * LPR + lane;
                if (j < A) {
                    float zj = prm.stage_z ? smem[L.z + (size_t)r * A + j] : (rawp[j] * scale - amp[j]);
                    float lp = zj - m - lsum;
                    float pj = expf(lp);
                    float dz = -w * ((j == act ? 1.0f : 0.0f) - pj);
                    if (pj > 0.0f) dz += k * pj * (lp + h);
                    st_stream(outp + j, dz * scale);
                }
            }
        }
        if (lane == 0) {
            if (prm.has_v) a.dvalue_raw[grow] = gv;
            if (prm.has_r) a.dreturn_raw[grow] = gr;
        }
    }
    // burn-in steps take no part in the loss: zero gradients (train.py:220-222)
    if (bi > 0) {
        const int nz = bi * Pa * A, nzr = bi * Pa;
        for (int e = 0; e < nE; e++) {
            size_t g0 = (size_t)(b0 + e) * T0 * Pa;
            for (int i = tid; i < nz; i += nthr) a.dpolicy_raw[g0 * A + i] = 0.0f;
            for (int i = tid; i < nzr; i += nthr) {
                if (prm.has_v) a.dvalue_raw[g0 + i] = 0.0f;
                if (prm.has_r) a.dreturn_raw[g0 + i] = 0.0f;
            }
        }
    }

    // ---------------- phase 4: deterministic reduction of the six scalars
    float part[6] = {Lp, Lv, Lr, Lent, Lreg, dcnt};
    const int warp = tid >> 5, wl = tid & 31, nwarp = (nthr + 31) >> 5;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float v = warp_sum(part[i]);
        if (wl == 0) smem[L.red + i * 32 + warp] = v;
    }
    __syncthreads();
    unsigned int *counter = reinterpret_cast<unsigned int *>(a.workspace);
    float *partials = reinterpret_cast<float *>(reinterpret_cast<char *>(a.workspace) + 256);
    __shared__ bool is_last;
    if (tid == 0) {
        for (int i = 0; i < 6; i++) {
            float v = 0.f;
            for (int w2 = 0; w2 < nwarp; w2++) v += smem[L.red + i * 32 + w2];
            partials[(size_t)blockIdx.x * 8 + i] = v;
        }
        __threadfence();
        unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int blk = tid; blk < (int)gridDim.x; blk += nthr) {
#pragma unroll
            for (int i = 0; i < 6; i++) acc[i] += (double)__ldcg(partials + (size_t)blk * 8 + i);
        }
        double *dred = reinterpret_cast<double *>(smem + L.red);  // 8*32 floats = 128 doubles >= 6*16
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double v = warp_sum_d(acc[i]);
            if (wl == 0) dred[i * 16 + warp] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double s[6];
            for (int i = 0; i < 6; i++) {
                s[i] = 0;
                for (int w2 = 0; w2 < nwarp; w2++) s[i] += dred[i * 16 + w2];
            }
            double lv = 0.5 * s[1];
            a.losses[HRL_LOSS_P] = (float)s[0];
            a.losses[HRL_LOSS_V] = (float)lv;
            a.losses[HRL_LOSS_R] = (float)s[2];
            a.losses[HRL_LOSS_ENT] = (float)s[3];
            a.losses[HRL_LOSS_TOTAL] = (float)(s[0] + lv + s[2] - (double)creg * s[4]);  // train.py:211-213
            a.losses[HRL_LOSS_DCNT] = (float)s[5];
            *counter = 0u;  // leave the workspace ready for the next launch
        }
    }
}

template <int LPR, int NPL>
static int launch(const LossParams &prm, int grid, int threads, size_t smem_bytes, cudaStream_t stream) {
    auto kern = loss_fwd_bwd_kernel<LPR, NPL>;
    if (smem_bytes > 48 * 1024)
        HRL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    kern<<<grid, threads, smem_bytes, stream>>>(prm);
    HRL_CUDA_CHECK(cudaGetLastError());
    return HRL_OK;
}

static int pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace hrl

extern "C" size_t hrl_loss_workspace_bytes(int32_t B, int32_t, int32_t, int32_t, int32_t) {
    return 256 + (size_t)(B > 0 ? B : 0) * 8 * sizeof(float);
}

extern "C" int hrl_loss_fwd_bwd(const HrlLossArgs *args, void *stream_) {
    using namespace hrl;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HRL_REQUIRE(args != nullptr, HRL_ERR_BAD_ARG, "hrl_loss_fwd_bwd: args is NULL");
    const HrlLossArgs &a = *args;
    HRL_REQUIRE(a.B > 0 && a.T > 0 && a.P > 0 && a.A > 0, HRL_ERR_BAD_ARG,
                "hrl_loss_fwd_bwd: non-positive dimension (B=%d T=%d P=%d A=%d)", a.B, a.T, a.P, a.A);
    HRL_REQUIRE(a.Pa == 1 || a.Pa == a.P, HRL_ERR_BAD_ARG, "hrl_loss_fwd_bwd: Pa must be 1 or P (Pa=%d P=%d)", a.Pa, a.P);
    HRL_REQUIRE(a.burn_in >= 0 && a.burn_in < a.T, HRL_ERR_BAD_ARG, "hrl_loss_fwd_bwd: burn_in=%d outside [0,T=%d)", a.burn_in, a.T);
    HRL_REQUIRE(a.value_target >= 0 && a.value_target <= 3 && a.policy_target >= 0 && a.policy_target <= 3,
                HRL_ERR_BAD_ARG, "hrl_loss_fwd_bwd: unknown target algorithm (value=%d policy=%d)", a.value_target, a.policy_target);
    HRL_REQUIRE(a.policy_raw && a.action_mask && a.action && a.selected_prob && a.reward && a.ret && a.turn_mask &&
                    a.observation_mask && a.episode_mask && a.progress && a.outcome,
                HRL_ERR_BAD_ARG, "hrl_loss_fwd_bwd: a required input pointer is NULL");
    HRL_REQUIRE(a.dpolicy_raw && a.losses, HRL_ERR_BAD_ARG, "hrl_loss_fwd_bwd: a required output pointer is NULL");
    HRL_REQUIRE((a.value_raw != nullptr) == (a.dvalue_raw != nullptr) && (a.return_raw != nullptr) == (a.dreturn_raw != nullptr),
                HRL_ERR_BAD_ARG, "hrl_loss_fwd_bwd: each head needs both its output and its gradient buffer");
    HRL_REQUIRE(a.A <= 512, HRL_ERR_UNSUPPORTED, "hrl_loss_fwd_bwd: A=%d > 512 not built", a.A);
    HRL_REQUIRE(a.P <= 64, HRL_ERR_UNSUPPORTED, "hrl_loss_fwd_bwd: P=%d > 64 not built", a.P);

    LossParams prm;
    prm.a = a;
    prm.Tt = a.T - a.burn_in;
    prm.has_v = a.value_raw != nullptr;
    prm.has_r = a.return_raw != nullptr;

    const int LPR = pow2_ceil(a.A) < 32 ? pow2_ceil(a.A) : 32;
    const int NPL = pow2_ceil((a.A + LPR - 1) / LPR);
    const int R = prm.Tt * a.Pa;
    const long lanes = (long)R * LPR;
    int threads, EPB;
    if (lanes >= 128) {
        EPB = 1;
        threads = 128;
        while (threads * 2 <= lanes && threads < 512) threads *= 2;
    } else {
        threads = 128;
        EPB = (int)(128 / lanes);
        if (EPB < 1) EPB = 1;
        if (EPB > a.B) EPB = a.B;
    }
    if (threads < EPB * a.P) threads = ((EPB * a.P + 31) / 32) * 32;
    prm.EPB = EPB;
    const int grid = (a.B + EPB - 1) / EPB;

    int dev = 0, max_smem = 0;
    HRL_CUDA_CHECK(cudaGetDevice(&dev));
    HRL_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    prm.stage_z = 1;
    SmemLayout L = make_layout(EPB, prm.Tt, a.P, a.Pa, a.A, 1);
    if ((size_t)L.total * 4 > (size_t)max_smem - 1024) {
        prm.stage_z = 0;
        L = make_layout(EPB, prm.Tt, a.P, a.Pa, a.A, 0);
    }
    const size_t smem_bytes = (size_t)L.total * 4;
    HRL_REQUIRE(smem_bytes <= (size_t)max_smem - 1024, HRL_ERR_UNSUPPORTED,
                "hrl_loss_fwd_bwd: T=%d P=%d needs %zu bytes of shared memory (> %d)", a.T, a.P, smem_bytes, max_smem);
    HRL_REQUIRE(a.workspace != nullptr && a.workspace_bytes >= 256 + (size_t)grid * 8 * sizeof(float), HRL_ERR_WORKSPACE,
                "hrl_loss_fwd_bwd: workspace of %zu bytes is too small (need %zu)", a.workspace_bytes,
                256 + (size_t)grid * 8 * sizeof(float));

#define HRL_CASE(l, n) \
    if (LPR == l && NPL == n) return launch<l, n>(prm, grid, threads, smem_bytes, stream);
    HRL_CASE(1, 1) HRL_CASE(2, 1) HRL_CASE(4, 1) HRL_CASE(8, 1) HRL_CASE(16, 1) HRL_CASE(32, 1)
    HRL_CASE(32, 2) HRL_CASE(32, 4) HRL_CASE(32, 8) HRL_CASE(32, 16)
#undef HRL_CASE
    set_error("hrl_loss_fwd_bwd: no kernel for LPR=%d NPL=%d", LPR, NPL);
    return HRL_ERR_UNSUPPORTED;
}
