// Stand-alone compute_target (handyrl/losses.py:63-80): one thread per (b,p) column walks T in
// reverse.  Kept for API parity with the reference function; the learner itself uses the fused
// kernel in loss_kernel.cu, which runs the same recurrences on chip.
#include "common.cuh"

namespace hrl {

__global__ void __launch_bounds__(128) compute_target_kernel(
    int algo, int B, int T, int P, int Tr, int Pr, const float *__restrict__ values,
    const float *__restrict__ returns, const float *__restrict__ rewards, float lmb, float gamma,
    const float *__restrict__ rhos, const float *__restrict__ cs, const float *__restrict__ masks,
    float *__restrict__ targets, float *__restrict__ advantages) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B * P) return;
    const int b = c / P, p = c - b * P;
    const size_t base = (size_t)b * T * P + p;
    const size_t rbase = (size_t)b * Tr * P + p;
    const size_t hbase = (size_t)b * T * Pr + (Pr == 1 ? 0 : p);
    const float boot = returns[rbase + (size_t)(Tr - 1) * P];  // returns[:, -1]
    float G = 0.f, acc = 0.f, vs_next = 0.f, v_next = 0.f, lam_next = 0.f;
    for (int t = T - 1; t >= 0; t--) {
        const size_t i = base + (size_t)t * P;
        const float ret_t = returns[rbase + (size_t)(Tr == 1 ? 0 : t) * P];
        float tgt, adv;
        if (values == nullptr) {  // losses.py:64-66
            tgt = ret_t;
            adv = ret_t;
        } else {
            const float v = values[i];
            const float r = rewards ? rewards[i] : 0.0f;
            const float mk = masks ? masks[i] : 1.0f;
            const float lam = lmb + (1.0f - lmb) * (1.0f - mk);  // losses.py:71
            const bool last = (t == T - 1);
            if (algo == HRL_MC) {
                tgt = ret_t;
                adv = ret_t - v;
            } else if (algo == HRL_TD) {
                G = last ? boot : r + gamma * ((1.0f - lam_next) * v_next + lam_next * G);
                tgt = G;
                adv = G - v;
            } else if (algo == HRL_UPGO) {
                G = last ? boot : r + gamma * fmaxf(v_next, (1.0f - lam_next) * v_next + lam_next * G);
                tgt = G;
                adv = G - v;
            } else {
                const float rho = rhos ? rhos[hbase + (size_t)t * Pr] : 1.0f;
                const float cc = cs ? cs[hbase + (size_t)t * Pr] : 1.0f;
                const float delta = rho * (r + gamma * (last ? boot : v_next) - v);
                acc = last ? delta : delta + gamma * lam_next * cc * acc;
                tgt = acc + v;
                adv = r + gamma * (last ? boot : vs_next) - v;
                vs_next = tgt;
            }
            v_next = v;
            lam_next = lam;
        }
        targets[i] = tgt;
        advantages[i] = adv;
    }
}

}  // namespace hrl

extern "C" int hrl_compute_target(int32_t algo, int32_t B, int32_t T, int32_t P, int32_t Tr, int32_t Pr,
                                  const float *values, const float *returns, const float *rewards, float lambda,
                                  float gamma, const float *rhos, const float *cs, const float *masks, float *targets,
                                  float *advantages, void *stream) {
    using namespace hrl;
    HRL_REQUIRE(algo >= 0 && algo <= 3, HRL_ERR_BAD_ARG, "hrl_compute_target: no algorithm with id %d", algo);
    HRL_REQUIRE(B > 0 && T > 0 && P > 0, HRL_ERR_BAD_ARG, "hrl_compute_target: non-positive dimension");
    HRL_REQUIRE((Tr == T || Tr == 1) && (Pr == P || Pr == 1), HRL_ERR_BAD_ARG,
                "hrl_compute_target: returns/rhos broadcast dims must be 1 or full (Tr=%d Pr=%d)", Tr, Pr);
    HRL_REQUIRE(returns && targets && advantages, HRL_ERR_BAD_ARG, "hrl_compute_target: NULL returns/targets/advantages");
    const int threads = 128, grid = (B * P + threads - 1) / threads;
    compute_target_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        algo, B, T, P, Tr, Pr, values, returns, rewards, lambda, gamma, rhos, cs, masks, targets, advantages);
    HRL_CUDA_CHECK(cudaGetLastError());
    return HRL_OK;
}
