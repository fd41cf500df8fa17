// Device-side building blocks shared by the fused loss kernels (loss_kernel.cu).
#pragma once
#include "common.cuh"
#include <math.h>

namespace hrl {

struct LossParams {
    HrlLossArgs a;
    int Tt;       // trained steps = T - burn_in
    int EPB;      // episodes per CTA
    int stage_z;  // masked logits kept in shared memory between the statistics and the gradient phase
    int has_v, has_r;
    int row_stride;   // floats between consecutive rows of the staged logits (>= A)
    int n_stage, chunk_rows;   // bulk pipeline: action-mask ring depth and rows per chunk
    long long *trace;  // optional per-phase clock64 stamps of one CTA (HRL_LOSS_TRACE, debugging only)
};

// shared-memory carve-up, in floats
struct SmemLayout {
    int emask, prog;                                        // [cells]
    int tm, om, rew, ret, wterm, dv, dr;                    // [cols]
    int vb, lamv, rout, lamr, tgv, tgr, advv, advr;         // [cols] recurrence inputs / outputs
    int outcome;                                            // [EPB*P]
    int logp, rho, ent, mx, lsum, scale, vraw, rraw, prob;  // [rows]
    int act;                                                // [rows] int64 (2 floats each)
    int red;                                                // [8*32]
    int bars;                                               // mbarriers (bulk pipeline), 8-byte aligned
    int z;                                                  // [rows*row_stride] if staged
    int am;                                                 // staged action mask: [rows*row_stride] or ring
    int total;
};

enum { kMaxChunks = 64, kMaxStages = 8 };

__host__ __device__ inline SmemLayout make_layout(int EPB, int Tt, int P, int Pa, int stage_z, int row_stride,
                                                  int am_floats) {
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
    L.vb = o; o += cols;
    L.lamv = o; o += cols;
    L.rout = o; o += cols;
    L.lamr = o; o += cols;
    L.tgv = o; o += cols;
    L.tgr = o; o += cols;
    L.advv = o; o += cols;
    L.advr = o; o += cols;
    L.outcome = o; o += EPB * P;
    L.logp = o; o += rows;
    L.rho = o; o += rows;
    L.ent = o; o += rows;
    L.mx = o; o += rows;
    L.lsum = o; o += rows;
    L.scale = o; o += rows;
    L.vraw = o; o += rows;
    L.rraw = o; o += rows;
    L.prob = o; o += rows;
    o = (o + 1) & ~1;
    L.act = o; o += 2 * rows;
    o = (o + 3) & ~3;
    L.red = o; o += 8 * 32;
    L.bars = o; o += 2 * (kMaxChunks + 2 * kMaxStages);
    o = (o + 31) & ~31;   // 128-byte alignment for the bulk-copy destinations
    L.z = o;
    if (stage_z) o += rows * row_stride;
    o = (o + 31) & ~31;
    L.am = o;
    o += am_floats;
    L.total = o;
    return L;
}

// ---------------------------------------------------------------- async copy primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copies (SASS: UBLKCP); sizes and addresses are multiples of 16 bytes
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// 2^x on the SFU (MUFU.EX2, max relative error 2^-22); callers fold log2(e) into the argument with one FMA
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------- recurrences (losses.py:16-60)
// One reverse-time recurrence over a column of Tt steps held in shared memory (stride st floats).
// Only the loop-carried arithmetic lives here; inputs of step t-1 are fetched before the dependent
// math of step t so that shared-memory latency stays off the critical path.
__device__ __forceinline__ void run_chain(int algo, bool has_baseline, int Tt, int st, const float *__restrict__ v,
                                          const float *__restrict__ lam, const float *__restrict__ rew, bool rew_zero,
                                          const float *__restrict__ ret_all, float ret_const, bool ret_is_const,
                                          float gamma, float boot, const float *__restrict__ rho, int rho_st,
                                          float *__restrict__ tgt, float *__restrict__ adv) {
    if (!has_baseline || algo == HRL_MC) {  // losses.py:64-66, 16-17: no recurrence
        for (int t = 0; t < Tt; t++) {
            const float r = ret_is_const ? ret_const : ret_all[t * st];
            if (tgt) tgt[t * st] = r;
            adv[t * st] = has_baseline ? r - v[t * st] : r;
        }
        return;
    }
    if (algo == HRL_TD || algo == HRL_UPGO) {  // losses.py:20-42
        const bool up = (algo == HRL_UPGO);
        float G = boot;
        float v_next = v[(Tt - 1) * st], lam_next = lam[(Tt - 1) * st];
        if (tgt) tgt[(Tt - 1) * st] = G;
        adv[(Tt - 1) * st] = G - v_next;
        int tp = Tt >= 2 ? Tt - 2 : 0;
        float v_t = v[tp * st], lam_t = lam[tp * st], r_t = rew_zero ? 0.0f : rew[tp * st];
        for (int t = Tt - 2; t >= 0; t--) {
            tp = t >= 1 ? t - 1 : 0;
            const float v_p = v[tp * st], lam_p = lam[tp * st], r_p = rew_zero ? 0.0f : rew[tp * st];
            float mix = (1.0f - lam_next) * v_next + lam_next * G;
            if (up) mix = fmaxf(v_next, mix);
            G = r_t + gamma * mix;
            if (tgt) tgt[t * st] = G;
            adv[t * st] = G - v_t;
            v_next = v_t; lam_next = lam_t;
            v_t = v_p; lam_t = lam_p; r_t = r_p;
        }
        return;
    }
    // V-Trace, losses.py:45-60 (c-bar == rho-bar: both clip thresholds are 1, train.py:229, 237-238)
    float v_next = boot, vs_next = boot, acc = 0.0f, lam_next = 0.0f;
    int tp = Tt - 1;
    float v_t = v[tp * st], lam_t = lam[tp * st], r_t = rew_zero ? 0.0f : rew[tp * st], rh_t = rho[tp * rho_st];
    for (int t = Tt - 1; t >= 0; t--) {
        tp = t >= 1 ? t - 1 : 0;
        const float v_p = v[tp * st], lam_p = lam[tp * st], r_p = rew_zero ? 0.0f : rew[tp * st], rh_p = rho[tp * rho_st];
        const float delta = rh_t * (r_t + gamma * v_next - v_t);
        acc = (t == Tt - 1) ? delta : delta + gamma * lam_next * rh_t * acc;
        const float vs = acc + v_t;
        adv[t * st] = r_t + gamma * vs_next - v_t;
        if (tgt) tgt[t * st] = vs;
        vs_next = vs; v_next = v_t; lam_next = lam_t;
        v_t = v_p; lam_t = lam_p; r_t = r_p; rh_t = rh_p;
    }
}

struct CtaCtx {
    int T0, P, Pa, A, bi, Tt;
    int b0, nE, tid, nthr;
    int nrows, ncols, ncells;
};

// phases 2a/2b/2c: from per-row statistics (logp, rho, ent in smem) to per-cell gradient factors and the six
// loss partial sums of this thread.  Caller must __syncthreads() before (statistics visible) and after.
__device__ __forceinline__ void targets_and_losses(const LossParams &prm, const SmemLayout &L, float *smem, const CtaCtx &c,
                                                   float part[6]) {
    const HrlLossArgs &a = prm.a;
    const int P = c.P, Pa = c.Pa, Tt = c.Tt;
    // ---- 2a: everything the recurrences need that is local to a cell, in parallel
    const bool sym = a.two_player_zero_sum && P == 2;
    for (int i = c.tid; i < c.ncols; i += c.nthr) {
        const int cell = i / P, p = i - cell * P;
        const int q = (Pa == P) ? p : 0;
        const int e = cell / Tt;
        const float om = smem[L.om + i], em = smem[L.emask + cell];
        const float vout = smem[L.vraw + cell * Pa + q] * om;      // train.py:184
        float vb = vout, vm = om;
        if (sym) {  // train.py:243-247
            const int po = 1 - p, qo = (Pa == P) ? po : 0;
            const float omo = smem[L.om + cell * P + po];
            const float vo = -(smem[L.vraw + cell * Pa + qo] * omo);
            vb = (vout * om + vo * omo) / (om + omo + 1e-8f);
            vm = fminf(fmaxf(om + omo, 0.0f), 1.0f);
        }
        smem[L.vb + i] = vb * em + smem[L.outcome + e * P + p] * (1.0f - em);   // train.py:248
        smem[L.lamv + i] = a.lambda + (1.0f - a.lambda) * (1.0f - vm);          // losses.py:71
        smem[L.rout + i] = smem[L.rraw + cell * Pa + q] * om;
        smem[L.lamr + i] = a.lambda + (1.0f - a.lambda) * (1.0f - om);
    }
    __syncthreads();

    // ---- 2b: the recurrences.  job = (column, kind); kind k runs in warp k so the four kinds
    // (value/return stream x target/advantage algorithm) proceed concurrently:
    //   kind 0: value stream, value_target   -> tgv (+ advv when policy_target == value_target)
    //   kind 1: return stream, value_target  -> tgr (+ advr ...)
    //   kind 2: value stream, policy_target  -> advv      (train.py:260-262)
    //   kind 3: return stream, policy_target -> advr
    {
        const bool two = (a.policy_target != a.value_target);
        const int nkind = two ? 4 : 2;
        const int warp_id = c.tid >> 5, lane_id = c.tid & 31, nwarps = c.nthr >> 5;
        const int ncolumn = c.nE * P;
        for (int kind = warp_id; kind < nkind; kind += nwarps) {
            const bool ret_stream = (kind & 1), adv_only = (kind >= 2);
            const int algo = adv_only ? a.policy_target : a.value_target;
            for (int col = lane_id; col < ncolumn; col += 32) {
                const int e = col / P, p = col - e * P;
                const int q = (Pa == P) ? p : 0;
                const int base = e * Tt * P + p;
                float *tgt = adv_only ? nullptr : smem + (ret_stream ? L.tgr : L.tgv) + base;
                float *adv = smem + (ret_stream ? L.advr : L.advv) + base;
                // with two algorithms, kinds 0/1 only produce targets: park their advantages in dv/dr (overwritten in 2c)
                if (two && !adv_only) adv = smem + (ret_stream ? L.dr : L.dv) + base;
                const float *rho = smem + L.rho + e * Tt * Pa + q;
                if (!ret_stream) {
                    const float oc = smem[L.outcome + e * P + p];
                    run_chain(algo, prm.has_v, Tt, P, smem + L.vb + base, smem + L.lamv + base, nullptr, true, nullptr, oc,
                              true, 1.0f, oc, rho, Pa, tgt, adv);
                } else {
                    const float boot = smem[L.ret + (e * Tt + Tt - 1) * P + p];   // returns[:, -1]
                    run_chain(algo, prm.has_r, Tt, P, smem + L.rout + base, smem + L.lamr + base, smem + L.rew + base, false,
                              smem + L.ret + base, 0.0f, false, a.gamma, boot, rho, Pa, tgt, adv);
                }
            }
        }
    }
    __syncthreads();

    // ---- 2c: per-cell loss terms and gradient factors, in parallel
    float Lp = 0.f, Lv = 0.f, Lr = 0.f, Lent = 0.f, Lreg = 0.f, dcnt = 0.f;
    for (int i = c.tid; i < c.ncols; i += c.nthr) {
        const int cell = i / P, p = i - cell * P;
        const int q = (Pa == P) ? p : 0, row = cell * Pa + q;
        const float om = smem[L.om + i], tm = smem[L.tm + i];
        const float rho = smem[L.rho + row];
        const float tgv = smem[L.tgv + i], tgr = smem[L.tgr + i];
        const float tot_adv = rho * (smem[L.advv + i] + smem[L.advr + i]);     // train.py:265
        smem[L.wterm + i] = tot_adv * tm;
        Lp += -smem[L.logp + row] * tot_adv * tm;                               // train.py:202
        float dv = 0.f, dr = 0.f;
        if (prm.has_v) {                                                        // train.py:204
            const float d = smem[L.vraw + row] * om - tgv;
            Lv += d * d * om;
            dv = d * om * om;
        }
        if (prm.has_r) {                                                        // train.py:206 smooth_l1, beta 1
            const float d = smem[L.rout + i] - tgr, ad = fabsf(d);
            Lr += (ad < 1.0f ? 0.5f * d * d : ad - 0.5f) * om;
            dr = fminf(fmaxf(d, -1.0f), 1.0f) * om * om;
        }
        smem[L.dv + i] = dv;
        smem[L.dr + i] = dr;
        const float h = smem[L.ent + row] * tm;                                 // train.py:208
        Lent += h;
        Lreg += h * (1.0f - smem[L.prog + cell] * (1.0f - a.entropy_regularization_decay));   // train.py:212
        dcnt += tm;
        if (a.tap_target_value || a.tap_target_return || a.tap_advantage) {
            const int e = cell / Tt, t = cell - e * Tt;
            const size_t gcol = ((size_t)(c.b0 + e) * c.T0 + c.bi + t) * P + p;
            if (a.tap_target_value) a.tap_target_value[gcol] = tgv;
            if (a.tap_target_return) a.tap_target_return[gcol] = tgr;
            if (a.tap_advantage) a.tap_advantage[gcol] = tot_adv;
        }
    }
    part[0] = Lp; part[1] = Lv; part[2] = Lr; part[3] = Lent; part[4] = Lreg; part[5] = dcnt;
}

// Block-reduce the six partial sums and publish them; returns (block-uniform) whether this CTA was the last
// one to publish.  Called BEFORE the gradient phase so that the fence does not wait for the bulk of the stores.
__device__ __forceinline__ bool publish_partials(const LossParams &prm, const SmemLayout &L, float *smem, const CtaCtx &c,
                                                 const float part[6], bool *s_flag) {
    const int warp = c.tid >> 5, wl = c.tid & 31, nwarp = (c.nthr + 31) >> 5;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float v = warp_sum(part[i]);
        if (wl == 0) smem[L.red + i * 32 + warp] = v;
    }
    __syncthreads();
    unsigned int *counter = reinterpret_cast<unsigned int *>(prm.a.workspace);
    float *partials = reinterpret_cast<float *>(reinterpret_cast<char *>(prm.a.workspace) + 256);
    if (c.tid < 6) {
        float v = 0.f;
        for (int w2 = 0; w2 < nwarp; w2++) v += smem[L.red + c.tid * 32 + w2];
        partials[(size_t)blockIdx.x * 8 + c.tid] = v;
        __threadfence();
    }
    __syncwarp();
    if (c.tid == 0) {
        unsigned int ticket = atomicAdd(counter, 1u);
        *s_flag = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    return *s_flag;
}

// Executed by the last CTA only: fold all block partials in a fixed order (fp64) and write the scalars.
__device__ __forceinline__ void finalize_losses(const LossParams &prm, const SmemLayout &L, float *smem, const CtaCtx &c) {
    const int warp = c.tid >> 5, wl = c.tid & 31, nwarp = (c.nthr + 31) >> 5;
    unsigned int *counter = reinterpret_cast<unsigned int *>(prm.a.workspace);
    const float *partials = reinterpret_cast<const float *>(reinterpret_cast<const char *>(prm.a.workspace) + 256);
    __threadfence();
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int blk = c.tid; blk < (int)gridDim.x; blk += c.nthr) {
#pragma unroll
        for (int i = 0; i < 6; i++) acc[i] += (double)__ldcg(partials + (size_t)blk * 8 + i);
    }
    double *dred = reinterpret_cast<double *>(smem + L.red);  // 8*32 floats = 128 doubles >= 6*20
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double v = warp_sum_d(acc[i]);
        if (wl == 0) dred[i * 20 + warp] = v;
    }
    __syncthreads();
    if (c.tid == 0) {
        double s[6];
        for (int i = 0; i < 6; i++) {
            s[i] = 0;
            for (int w2 = 0; w2 < nwarp; w2++) s[i] += dred[i * 20 + w2];
        }
        const double lv = 0.5 * s[1];
        float *out = prm.a.losses;
        out[HRL_LOSS_P] = (float)s[0];
        out[HRL_LOSS_V] = (float)lv;
        out[HRL_LOSS_R] = (float)s[2];
        out[HRL_LOSS_ENT] = (float)s[3];
        out[HRL_LOSS_TOTAL] = (float)(s[0] + lv + s[2] - (double)prm.a.entropy_regularization * s[4]);  // train.py:211-213
        out[HRL_LOSS_DCNT] = (float)s[5];
        *counter = 0u;  // leave the workspace ready for the next launch
    }
}

// per-row gradient factors gathered from the per-cell terms (sum over players when Pa == 1)
struct RowFactors {
    float w, k, gv, gr;
};
__device__ __forceinline__ RowFactors row_factors(const LossParams &prm, const SmemLayout &L, const float *smem, int cell, int q,
                                                  int P, int Pa) {
    RowFactors f = {0.f, 0.f, 0.f, 0.f};
    if (Pa == P) {
        f.w = smem[L.wterm + cell * P + q];
        f.k = smem[L.tm + cell * P + q];
        f.gv = smem[L.dv + cell * P + q];
        f.gr = smem[L.dr + cell * P + q];
    } else {
        for (int p = 0; p < P; p++) {
            f.w += smem[L.wterm + cell * P + p];
            f.k += smem[L.tm + cell * P + p];
            f.gv += smem[L.dv + cell * P + p];
            f.gr += smem[L.dr + cell * P + p];
        }
    }
    f.w *= smem[L.emask + cell];
    f.k *= prm.a.entropy_regularization * (1.0f - smem[L.prog + cell] * (1.0f - prm.a.entropy_regularization_decay));
    return f;
}

// dL/dz_j for one element (closed form), times d z / d raw = scale
__device__ __forceinline__ float grad_elem(float zj, bool is_act, float m, float lsum, float h, float w, float k, float scale) {
    const float lp = zj - m - lsum;
    const float pj = expf(lp);
    float dz = -w * ((is_act ? 1.0f : 0.0f) - pj);
    if (pj > 0.0f) dz += k * pj * (lp + h);
    return dz * scale;
}

}  // namespace hrl
