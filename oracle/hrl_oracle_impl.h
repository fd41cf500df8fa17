/*
 * TEST INFRASTRUCTURE -- CPU restatement of the reference algorithm, not product code.
 * Included twice by hrl_oracle.c with REAL = float / double and SUF = _f32 / _f64.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks every function here against
 * the tests/golden fixtures, which are outputs of the reference itself (tests/golden/gen_golden.py).
 *
 * Each block cites the reference lines it restates (paths relative to DeNA/HandyRL).
 * The structure is deliberately "one tensor op = one loop" in the reference's order.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ---- handyrl/losses.py:63-80 compute_target and the four algorithms, one (b,p) column ---- */
/* v, rew, lam, rho, c: arrays over t with stride 1 (already gathered); returns_last: returns[:, -1] */
static void FN(column_target)(int algo, int T, const REAL *v, const REAL *returns_all, REAL returns_last,
                              const REAL *rew, const REAL *mask, REAL lmb, REAL gamma,
                              const REAL *rho, const REAL *c, REAL *tgt, REAL *adv)
{
    int t;
    if (v == NULL) { /* losses.py:64-66 */
        for (t = 0; t < T; t++) { tgt[t] = returns_all[t]; adv[t] = returns_all[t]; }
        return;
    }
    if (algo == HRL_MC) { /* losses.py:16-17 */
        for (t = 0; t < T; t++) { tgt[t] = returns_all[t]; adv[t] = returns_all[t] - v[t]; }
        return;
    }
    REAL *lam = (REAL *)malloc(sizeof(REAL) * (size_t)T);
    for (t = 0; t < T; t++) lam[t] = lmb + (1 - lmb) * (1 - mask[t]); /* losses.py:71 */

    if (algo == HRL_TD) { /* losses.py:20-29 */
        tgt[T - 1] = returns_last;
        for (t = T - 2; t >= 0; t--)
            tgt[t] = rew[t] + gamma * ((1 - lam[t + 1]) * v[t + 1] + lam[t + 1] * tgt[t + 1]);
        for (t = 0; t < T; t++) adv[t] = tgt[t] - v[t];
    } else if (algo == HRL_UPGO) { /* losses.py:32-42 */
        tgt[T - 1] = returns_last;
        for (t = T - 2; t >= 0; t--) {
            REAL mix = (1 - lam[t + 1]) * v[t + 1] + lam[t + 1] * tgt[t + 1];
            tgt[t] = rew[t] + gamma * (v[t + 1] > mix ? v[t + 1] : mix);
        }
        for (t = 0; t < T; t++) adv[t] = tgt[t] - v[t];
    } else { /* HRL_VTRACE, losses.py:45-60 */
        REAL *acc = (REAL *)malloc(sizeof(REAL) * (size_t)T);
        for (t = 0; t < T; t++) {
            REAL vnext = (t == T - 1) ? returns_last : v[t + 1];
            acc[t] = rho[t] * (rew[t] + gamma * vnext - v[t]); /* deltas, :48 */
        }
        for (t = T - 2; t >= 0; t--) acc[t] = acc[t] + gamma * lam[t + 1] * c[t] * acc[t + 1]; /* :53 */
        for (t = 0; t < T; t++) tgt[t] = acc[t] + v[t]; /* vs, :56 */
        for (t = 0; t < T; t++) {
            REAL vsnext = (t == T - 1) ? returns_last : tgt[t + 1];
            adv[t] = rew[t] + gamma * vsnext - v[t]; /* :57-58 */
        }
        free(acc);
    }
    free(lam);
}

/* ---- stand-alone compute_target on (B,T,P) tensors with the reference's broadcasting ---- */
int FN(hrl_oracle_compute_target)(int algo, int B, int T, int P, int Tr, int Pr,
                                  const float *values, const float *returns, const float *rewards,
                                  double lmb, double gamma, const float *rhos, const float *cs,
                                  const float *masks, REAL *targets, REAL *advantages)
{
    int b, p, t;
    REAL *buf = (REAL *)malloc(sizeof(REAL) * (size_t)T * 8);
    REAL *v = buf, *ra = buf + T, *rw = buf + 2 * T, *mk = buf + 3 * T, *rh = buf + 4 * T, *cc = buf + 5 * T,
         *tg = buf + 6 * T, *ad = buf + 7 * T;
    if (algo < 0 || algo > 3) { free(buf); return -1; }
    for (b = 0; b < B; b++)
        for (p = 0; p < P; p++) {
            for (t = 0; t < T; t++) {
                size_t i = ((size_t)b * T + t) * P + p;
                if (values) v[t] = values[i];
                ra[t] = returns[((size_t)b * Tr + (Tr == 1 ? 0 : t)) * P + p];
                rw[t] = rewards ? rewards[i] : 0;
                mk[t] = masks ? masks[i] : 1;
                rh[t] = rhos ? rhos[((size_t)b * T + t) * Pr + (Pr == 1 ? 0 : p)] : 1;
                cc[t] = cs ? cs[((size_t)b * T + t) * Pr + (Pr == 1 ? 0 : p)] : 1;
            }
            FN(column_target)(algo, T, values ? v : NULL, ra, ra[T - 1], rw, mk, (REAL)lmb, (REAL)gamma, rh, cc, tg, ad);
            for (t = 0; t < T; t++) {
                size_t i = ((size_t)b * T + t) * P + p;
                targets[i] = tg[t];
                advantages[i] = ad[t];
            }
        }
    free(buf);
    return 0;
}

/*
 * ---- the fused loss: forward_prediction's mask epilogue (train.py:176-184),
 *      compute_loss (train.py:218-267), compose_losses (train.py:189-215) and the
 *      gradient of `total` w.r.t. the raw net outputs (what train.py:369 backpropagates
 *      into the net).  Inputs are the fp32 device-layout arrays of HrlLossArgs, here on
 *      the host; outputs are REAL. ----
 */
typedef struct FN(OracleLossOut) {
    REAL *dpolicy_raw, *dvalue_raw, *dreturn_raw; /* same shapes as the inputs */
    REAL *losses;                                 /* [HRL_NUM_LOSS] */
    REAL *target_value, *target_return, *advantage, *logp, *rho, *entropy; /* optional taps */
} FN(OracleLossOut);

int FN(hrl_oracle_loss)(const HrlLossArgs *a, const FN(OracleLossOut) *o)
{
    const int B = a->B, T0 = a->T, P = a->P, Pa = a->Pa, A = a->A, bi = a->burn_in;
    const int T = T0 - bi; /* train.py:220-222: everything below sees only t >= burn_in */
    int b, t, p, q, j;
    if (B <= 0 || T <= 0 || P <= 0 || A <= 0 || (Pa != 1 && Pa != P)) return -1;
    if (a->value_target < 0 || a->value_target > 3 || a->policy_target < 0 || a->policy_target > 3) return -1;
    const int has_v = a->value_raw != NULL, has_r = a->return_raw != NULL;
    const REAL lmb = a->lambda, gamma = a->gamma;
    const REAL creg = a->entropy_regularization, cdec = a->entropy_regularization_decay;

    size_t nrow = (size_t)B * T * Pa, ncol = (size_t)B * T * P;
    REAL *z = (REAL *)malloc(sizeof(REAL) * nrow * A);     /* outputs['policy'] */
    REAL *prob = (REAL *)malloc(sizeof(REAL) * nrow * A);  /* softmax(policy) */
    REAL *scale = (REAL *)malloc(sizeof(REAL) * nrow);     /* d policy / d raw */
    REAL *logp = (REAL *)malloc(sizeof(REAL) * nrow);
    REAL *rho = (REAL *)malloc(sizeof(REAL) * nrow);
    REAL *ent = (REAL *)malloc(sizeof(REAL) * nrow);
    REAL *vout = (REAL *)calloc(ncol, sizeof(REAL)), *rout = (REAL *)calloc(ncol, sizeof(REAL));
    REAL *vbase = (REAL *)calloc(ncol, sizeof(REAL)), *vmask = (REAL *)calloc(ncol, sizeof(REAL));
    REAL *tgt_v = (REAL *)calloc(ncol, sizeof(REAL)), *adv_v = (REAL *)calloc(ncol, sizeof(REAL));
    REAL *tgt_r = (REAL *)calloc(ncol, sizeof(REAL)), *adv_r = (REAL *)calloc(ncol, sizeof(REAL));
    REAL *tot_adv = (REAL *)calloc(ncol, sizeof(REAL));
    REAL *col = (REAL *)malloc(sizeof(REAL) * (size_t)T * 10);

#define IN_ROW(b, t, q) ((((size_t)(b) * T0 + (t) + bi) * Pa) + (q))  /* index into (B,T0,Pa) inputs */
#define IN_COL(b, t, p) ((((size_t)(b) * T0 + (t) + bi) * P) + (p))   /* index into (B,T0,P) inputs  */
#define IN_CELL(b, t) ((size_t)(b) * T0 + (t) + bi)
#define ROW(b, t, q) ((((size_t)(b) * T + (t)) * Pa) + (q))
#define COL(b, t, p) ((((size_t)(b) * T + (t)) * P) + (p))

    /* train.py:176-181: policy = (raw * turn_mask)[summed over players when Pa == 1] - action_mask */
    for (b = 0; b < B; b++) for (t = 0; t < T; t++) for (q = 0; q < Pa; q++) {
        REAL s = 0;
        if (Pa == P) s = a->turn_mask[IN_COL(b, t, q)];
        else for (p = 0; p < P; p++) s += a->turn_mask[IN_COL(b, t, p)];
        scale[ROW(b, t, q)] = s;
        for (j = 0; j < A; j++) {
            REAL raw = a->policy_raw[IN_ROW(b, t, q) * A + j];
            z[ROW(b, t, q) * A + j] = raw * s - (REAL)a->action_mask[IN_ROW(b, t, q) * A + j];
        }
    }
    /* train.py:182-184: other heads are multiplied by observation_mask (raw has Pa columns) */
    for (b = 0; b < B; b++) for (t = 0; t < T; t++) for (p = 0; p < P; p++) {
        q = (Pa == P) ? p : 0;
        REAL om = a->observation_mask[IN_COL(b, t, p)];
        if (has_v) vout[COL(b, t, p)] = (REAL)a->value_raw[IN_ROW(b, t, q)] * om;
        if (has_r) rout[COL(b, t, p)] = (REAL)a->return_raw[IN_ROW(b, t, q)] * om;
    }

    /* train.py:231-238 and :208 -- log-softmax, gather, importance ratio, entropy */
    for (b = 0; b < B; b++) for (t = 0; t < T; t++) for (q = 0; q < Pa; q++) {
        size_t r = ROW(b, t, q);
        const REAL *zr = z + r * A;
        REAL m = zr[0], se = 0, h = 0;
        for (j = 1; j < A; j++) if (zr[j] > m) m = zr[j];
        for (j = 0; j < A; j++) se += MATH(exp)(zr[j] - m);
        REAL lse = MATH(log)(se);
        for (j = 0; j < A; j++) {
            REAL lp = zr[j] - m - lse;                 /* log_softmax */
            REAL pj = MATH(exp)(lp);
            prob[r * A + j] = pj;
            /* Categorical.entropy: -sum(p * clamp(logp, min=finfo.min)) */
            if (lp < -REAL_MAX) lp = -REAL_MAX;
            h -= pj * lp;
        }
        ent[r] = h;
        REAL e = a->episode_mask[IN_CELL(b, t)];
        int64_t act = a->action[IN_ROW(b, t, q)];
        REAL lt = (zr[act] - m - lse) * e;                                   /* :232 */
        REAL mu = a->selected_prob[IN_ROW(b, t, q)];
        if (mu < (REAL)1e-16) mu = (REAL)1e-16;
        if (mu > 1) mu = 1;
        REAL lb = MATH(log)(mu) * e;                                         /* :231 */
        REAL rr = MATH(exp)(lt - lb);                                        /* :235-236 */
        if (rr < 0) rr = 0;
        if (rr > 1) rr = 1;                                                  /* :237-238, both thresholds 1 */
        logp[r] = lt;
        rho[r] = rr;
    }

    /* train.py:241-248 -- value baseline: 2-player symmetrisation, then outcome past the end */
    for (b = 0; b < B; b++) for (t = 0; t < T; t++) for (p = 0; p < P; p++) {
        REAL om = a->observation_mask[IN_COL(b, t, p)];
        REAL vb = vout[COL(b, t, p)], vm = om;
        if (a->two_player_zero_sum && P == 2) {
            REAL omo = a->observation_mask[IN_COL(b, t, 1 - p)];
            REAL vo = -vout[COL(b, t, 1 - p)];
            vb = (vb * om + vo * omo) / (om + omo + (REAL)1e-8);
            vm = om + omo;
            if (vm < 0) vm = 0;
            if (vm > 1) vm = 1;
        }
        REAL e = a->episode_mask[IN_CELL(b, t)];
        vbase[COL(b, t, p)] = vb * e + (REAL)a->outcome[(size_t)b * P + p] * (1 - e);
        vmask[COL(b, t, p)] = vm;
    }

    /* train.py:254-262 -- targets by value_target, advantages by policy_target, two streams */
    for (b = 0; b < B; b++) for (p = 0; p < P; p++) {
        REAL *v = col, *ra = col + T, *rw = col + 2 * T, *mk = col + 3 * T, *rh = col + 4 * T,
             *tg = col + 5 * T, *ad = col + 6 * T, *tg2 = col + 7 * T;
        q = (Pa == P) ? p : 0;
        for (t = 0; t < T; t++) rh[t] = rho[ROW(b, t, q)];
        /* value stream: returns = outcome (B,1,P,1), rewards None, gamma 1 (:254) */
        for (t = 0; t < T; t++) {
            v[t] = vbase[COL(b, t, p)];
            ra[t] = a->outcome[(size_t)b * P + p];
            rw[t] = 0;
            mk[t] = vmask[COL(b, t, p)];
        }
        FN(column_target)(a->value_target, T, has_v ? v : NULL, ra, ra[T - 1], rw, mk, lmb, 1, rh, rh, tg, ad);
        if (a->policy_target != a->value_target)
            FN(column_target)(a->policy_target, T, has_v ? v : NULL, ra, ra[T - 1], rw, mk, lmb, 1, rh, rh, tg2, ad);
        for (t = 0; t < T; t++) { tgt_v[COL(b, t, p)] = tg[t]; adv_v[COL(b, t, p)] = ad[t]; }
        /* return stream (:255) */
        for (t = 0; t < T; t++) {
            v[t] = rout[COL(b, t, p)];
            ra[t] = a->ret[IN_COL(b, t, p)];
            rw[t] = a->reward[IN_COL(b, t, p)];
            mk[t] = a->observation_mask[IN_COL(b, t, p)];
        }
        FN(column_target)(a->value_target, T, has_r ? v : NULL, ra, ra[T - 1], rw, mk, lmb, gamma, rh, rh, tg, ad);
        if (a->policy_target != a->value_target)
            FN(column_target)(a->policy_target, T, has_r ? v : NULL, ra, ra[T - 1], rw, mk, lmb, gamma, rh, rh, tg2, ad);
        for (t = 0; t < T; t++) { tgt_r[COL(b, t, p)] = tg[t]; adv_r[COL(b, t, p)] = ad[t]; }
        /* :265 total_advantages = clipped_rhos * (adv_value + adv_return) */
        for (t = 0; t < T; t++) tot_adv[COL(b, t, p)] = rh[t] * (adv_v[COL(b, t, p)] + adv_r[COL(b, t, p)]);
    }

    /* train.py:189-215 compose_losses (sums) */
    REAL Lp = 0, Lv = 0, Lr = 0, Lent = 0, Lentreg = 0, dcnt = 0;
    for (b = 0; b < B; b++) for (t = 0; t < T; t++) for (p = 0; p < P; p++) {
        q = (Pa == P) ? p : 0;
        REAL tm = a->turn_mask[IN_COL(b, t, p)], om = a->observation_mask[IN_COL(b, t, p)];
        dcnt += tm;                                                                       /* :200 */
        Lp += -logp[ROW(b, t, q)] * tot_adv[COL(b, t, p)] * tm;                           /* :202 */
        if (has_v) { REAL d = vout[COL(b, t, p)] - tgt_v[COL(b, t, p)]; Lv += d * d * om; } /* :204 */
        if (has_r) {                                                                      /* :206 */
            REAL d = rout[COL(b, t, p)] - tgt_r[COL(b, t, p)], ad = d < 0 ? -d : d;
            Lr += (ad < 1 ? (REAL)0.5 * d * d : ad - (REAL)0.5) * om;
        }
        REAL h = ent[ROW(b, t, q)] * tm;                                                  /* :208 */
        Lent += h;                                                                        /* :209 */
        Lentreg += h * (1 - (REAL)a->progress[IN_CELL(b, t)] * (1 - cdec));              /* :212 */
    }
    Lv /= 2;
    o->losses[HRL_LOSS_P] = Lp;
    o->losses[HRL_LOSS_V] = has_v ? Lv : 0;
    o->losses[HRL_LOSS_R] = has_r ? Lr : 0;
    o->losses[HRL_LOSS_ENT] = Lent;
    o->losses[HRL_LOSS_TOTAL] = Lp + (has_v ? Lv : 0) + (has_r ? Lr : 0) + Lentreg * -creg;  /* :211-213 */
    o->losses[HRL_LOSS_DCNT] = dcnt;

    /* gradient of `total` w.r.t. the raw outputs; targets, advantages and rho are constants
     * in the reference graph (train.py:235, 239 detach) */
    for (size_t i = 0; i < (size_t)B * T0 * Pa * A; i++) o->dpolicy_raw[i] = 0;
    if (has_v) for (size_t i = 0; i < (size_t)B * T0 * Pa; i++) o->dvalue_raw[i] = 0;
    if (has_r) for (size_t i = 0; i < (size_t)B * T0 * Pa; i++) o->dreturn_raw[i] = 0;
    for (b = 0; b < B; b++) for (t = 0; t < T; t++) for (q = 0; q < Pa; q++) {
        size_t r = ROW(b, t, q);
        REAL e = a->episode_mask[IN_CELL(b, t)];
        REAL w = 0, k = 0;
        for (p = 0; p < P; p++) {
            if (Pa == P && p != q) continue;
            REAL tm = a->turn_mask[IN_COL(b, t, p)];
            w += tot_adv[COL(b, t, p)] * tm;
            k += tm;
        }
        w *= e;
        k *= creg * (1 - (REAL)a->progress[IN_CELL(b, t)] * (1 - cdec));
        int64_t act = a->action[IN_ROW(b, t, q)];
        for (j = 0; j < A; j++) {
            REAL pj = prob[r * A + j];
            REAL lp = pj > 0 ? MATH(log)(pj) : 0;
            REAL dz = -w * ((j == act ? 1 : 0) - pj) + k * pj * (lp + ent[r]);
            o->dpolicy_raw[IN_ROW(b, t, q) * A + j] = dz * scale[r];
        }
        for (p = 0; p < P; p++) {
            if (Pa == P && p != q) continue;
            REAL om = a->observation_mask[IN_COL(b, t, p)];
            if (has_v) o->dvalue_raw[IN_ROW(b, t, q)] += (vout[COL(b, t, p)] - tgt_v[COL(b, t, p)]) * om * om;
            if (has_r) {
                REAL d = rout[COL(b, t, p)] - tgt_r[COL(b, t, p)];
                REAL g = d < -1 ? -1 : (d > 1 ? 1 : d);
                o->dreturn_raw[IN_ROW(b, t, q)] += g * om * om;
            }
        }
    }

    /* taps in the full (B,T0,...) layout, zero for burn-in steps */
    if (o->target_value) for (size_t i = 0; i < (size_t)B * T0 * P; i++) o->target_value[i] = 0;
    if (o->target_return) for (size_t i = 0; i < (size_t)B * T0 * P; i++) o->target_return[i] = 0;
    if (o->advantage) for (size_t i = 0; i < (size_t)B * T0 * P; i++) o->advantage[i] = 0;
    if (o->logp) for (size_t i = 0; i < (size_t)B * T0 * Pa; i++) o->logp[i] = 0;
    if (o->rho) for (size_t i = 0; i < (size_t)B * T0 * Pa; i++) o->rho[i] = 0;
    if (o->entropy) for (size_t i = 0; i < (size_t)B * T0 * Pa; i++) o->entropy[i] = 0;
    for (b = 0; b < B; b++) for (t = 0; t < T; t++) {
        for (p = 0; p < P; p++) {
            if (o->target_value) o->target_value[IN_COL(b, t, p)] = tgt_v[COL(b, t, p)];
            if (o->target_return) o->target_return[IN_COL(b, t, p)] = tgt_r[COL(b, t, p)];
            if (o->advantage) o->advantage[IN_COL(b, t, p)] = tot_adv[COL(b, t, p)];
        }
        for (q = 0; q < Pa; q++) {
            if (o->logp) o->logp[IN_ROW(b, t, q)] = logp[ROW(b, t, q)];
            if (o->rho) o->rho[IN_ROW(b, t, q)] = rho[ROW(b, t, q)];
            if (o->entropy) o->entropy[IN_ROW(b, t, q)] = ent[ROW(b, t, q)];
        }
    }
#undef IN_ROW
#undef IN_COL
#undef IN_CELL
#undef ROW
#undef COL
    free(z); free(prob); free(scale); free(logp); free(rho); free(ent); free(vout); free(rout);
    free(vbase); free(vmask); free(tgt_v); free(adv_v); free(tgt_r); free(adv_r); free(tot_adv); free(col);
    return 0;
}

#undef FN
#undef CAT
#undef CAT_
