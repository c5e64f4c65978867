/* CPU restatement (plain C) of the reference's hottest loop, used ONLY as the timed CPU baseline in
 * bench.py (`cpu_baseline`, `--impl reference`) and cross-checked against the numpy oracle in tests/.
 * TEST / BENCH INFRASTRUCTURE -- never linked into the product library.
 *
 * Follows Training/loss_functions.jl of the reference loop for loop:
 *   kron_conj2                 :193-200   (second argument fastest)
 *   kron_scaleadd_KLD! (bulk)  :248-262   one fused pass per sample: yhat += bt*phi; k += kprev/scale; kprev = phi
 *   Loss_Grad_KLD              :322-379   per class, samples in order, final flush :367
 *   update_caches!             Training/RealRealHighDimension.jl:107-144
 * The reference runs this loop on ONE thread (@turbo SIMD).  `nthreads` > 1 splits the samples of a
 * class over pthreads with per-thread accumulators (a data-parallel variant the reference does
 * not have) so that the baseline may use every host core.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void kron2(const double* x1, int l1, const double* x2, int l2, double* out) {
    for (int i = 0; i < l1; i++)
        for (int j = 0; j < l2; j++) out[j + l2 * i] = x1[i] * x2[j];
}

typedef struct {
    const double *bt, *L, *R, *xl, *xr;
    int64_t b, e;
    int d, chi_l, chi_r;
    double* k;          /* out: accumulated phi~/yhat over [b, e) (flushed) */
    double loss;        /* out: sum -log(yhat^2) */
} kld_job;

/* the reference's per-sample pass, sample-sequential inside the job (loss_functions.jl:248-262, 353-367) */
static void* kld_worker(void* arg) {
    kld_job* J = (kld_job*)arg;
    const int d = J->d, la = d * J->chi_r, lb = d * J->chi_l;
    const int64_t D = (int64_t)la * lb;
    double* k = J->k;
    double* kprev = (double*)calloc(D, sizeof(double));
    double* xa = (double*)malloc(sizeof(double) * la);
    double* xb = (double*)malloc(sizeof(double) * lb);
    double yhat = 1.0, loss = 0.0;
    memset(k, 0, sizeof(double) * D);
    for (int64_t s = J->b; s < J->e; s++) {
        kron2(J->R + s * J->chi_r, J->chi_r, J->xr + s * d, d, xa);      /* xa = kron_conj2(REP, ps[rid]) */
        kron2(J->L + s * J->chi_l, J->chi_l, J->xl + s * d, d, xb);      /* xb = kron_conj2(LEP, ps[lid]) */
        const double scale = yhat;
        double y = 0.0;
        for (int i = 0; i < la; i++) {
            const double a = xa[i];
            const double* btr = J->bt + (int64_t)lb * i;
            double* kr = k + (int64_t)lb * i;
            double* kp = kprev + (int64_t)lb * i;
            for (int j = 0; j < lb; j++) {
                const double phi = a * xb[j];
                y += btr[j] * phi;
                kr[j] += kp[j] / scale;
                kp[j] = phi;
            }
        }
        yhat = y;
        loss += -log(y * y);                                              /* KLD_iter! :318 */
    }
    if (J->e > J->b)
        for (int64_t t = 0; t < D; t++) k[t] += kprev[t] / yhat;         /* final flush :367 */
    J->loss = loss;
    free(kprev); free(xa); free(xb);
    return NULL;
}

/* B: D x C column-major; L: N x chi_l row-major; R: N x chi_r; xl, xr: N x d; grad: D x C column-major.
 * returns the loss. */
double loss_grad_kld_ref(const double* B, const double* L, const double* R, const double* xl, const double* xr,
                         int64_t N, int d, int chi_l, int chi_r, const int64_t* counts, int C, int train_sep,
                         int nthreads, double* grad) {
    const int64_t D = (int64_t)d * chi_r * d * chi_l;
    double losses = 0.0;
    int64_t i0 = 0;
    if (nthreads < 1) nthreads = 1;
    kld_job* jobs = (kld_job*)malloc(sizeof(kld_job) * nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    double* kbuf = (double*)malloc(sizeof(double) * D * nthreads);
    for (int ci = 0; ci < C; ci++) {
        const int64_t cn = counts[ci];
        double* gk = grad + D * ci;
        for (int t = 0; t < nthreads; t++) {
            kld_job* J = &jobs[t];
            J->bt = B + D * ci; J->L = L; J->R = R; J->xl = xl; J->xr = xr;
            J->b = i0 + cn * t / nthreads; J->e = i0 + cn * (t + 1) / nthreads;
            J->d = d; J->chi_l = chi_l; J->chi_r = chi_r; J->k = kbuf + D * t; J->loss = 0.0;
            if (t > 0) pthread_create(&th[t], NULL, kld_worker, J);
        }
        kld_worker(&jobs[0]);
        for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
        double class_loss = 0.0;
        const double denom = train_sep ? (double)cn : (double)N;
        for (int64_t e = 0; e < D; e++) {
            double s = 0.0;
            for (int t = 0; t < nthreads; t++) s += kbuf[D * t + e];
            gk[e] = -s / denom;
        }
        for (int t = 0; t < nthreads; t++) class_loss += jobs[t].loss;
        losses += train_sep ? class_loss / cn : class_loss;
        i0 += cn;
    }
    free(jobs); free(th); free(kbuf);
    if (!train_sep) losses /= (double)N;
    return losses;
}

typedef struct {
    const double *x, *env, *W;
    double* out;
    int64_t b, e;
    int d, chi, chi_new;
} env_job;

static void* env_worker(void* arg) {
    env_job* J = (env_job*)arg;
    const int d = J->d, chi = J->chi;
    for (int64_t i = J->b; i < J->e; i++)
        for (int k = 0; k < J->chi_new; k++) {
            double acc = 0.0;
            for (int a = 0; a < chi; a++) {
                double t = 0.0;
                const double* w = J->W + (int64_t)d * (a + (int64_t)chi * k);
                for (int s = 0; s < d; s++) t += J->x[i * d + s] * w[s];
                acc += t * J->env[i * chi + a];
            }
            J->out[i * J->chi_new + k] = acc;
        }
    return NULL;
}

/* env_new[i][k] = sum_{s,a} x[i][s] * W[s + d*(a + chi*k)] * env[i][a]   (update_caches! :129,:139) */
void env_update_ref(const double* x, const double* env, const double* W, int64_t N, int d, int chi, int chi_new,
                    int nthreads, double* out) {
    if (nthreads < 1) nthreads = 1;
    env_job* jobs = (env_job*)malloc(sizeof(env_job) * nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        env_job* J = &jobs[t];
        J->x = x; J->env = env; J->W = W; J->out = out; J->d = d; J->chi = chi; J->chi_new = chi_new;
        J->b = N * t / nthreads; J->e = N * (t + 1) / nthreads;
        if (t > 0) pthread_create(&th[t], NULL, env_worker, J);
    }
    env_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    free(jobs); free(th);
}
