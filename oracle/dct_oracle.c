/*
 * dct_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the Deep Co-Training consistency hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (the CUDA library behind include/dct_b200.h)
 * never does.  Every function restates a piece of the reference, cited as
 * file:line into /root/reference (jizongFox/Deep-Co-Training-for-Semi-Supervised-
 * Image-Segmentation @ 431a60cd).  The reference is pure Python on top of
 * PyTorch ATen + NumPy (both un-vendored, un-pinned by the reference; the
 * container has torch 2.11.0+cu128 / numpy 2.3): the floating-point pieces are
 * restated op by op in the reference's order, the integer pieces exactly.
 *
 * PARITY PINNING: the reference's own tests hold no golden vectors for this path
 * (SURVEY.md section 8c), so the oracle is pinned against outputs of the
 * reference itself, generated in the build container by oracle/make_golden.py
 * (which imports /root/reference through oracle/ref_shim.py) and committed
 * under tests/golden/.  tests/test_oracle_vs_golden.py checks every function
 * here against those fixtures.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -mfma -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int dcto_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void dcto_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------- floating-point pieces: float and double twins ---------------- */
#define REAL float
#define FN(name) name##_f32
#define EXP expf
#define LOG logf
#define R(x) ((float)(x))
#include "dct_oracle_body.inc"
#undef REAL
#undef FN
#undef EXP
#undef LOG
#undef R

#define REAL double
#define FN(name) name##_f64
#define EXP exp
#define LOG log
#define R(x) ((double)(x))
#include "dct_oracle_body.inc"
#undef REAL
#undef FN
#undef EXP
#undef LOG
#undef R

/* ---------------- integer pieces ---------------- */

/*
 * The prediction map of DiceMeter.add (generalframework/metrics/dice_meter.py:25-29,50-55):
 *     pred = probs2class(F.softmax(pred_logit, 1))      utils/utils.py:178-184  (argmax, first index on ties)
 * argmax-after-softmax differs from argmax-of-logits only where two distinct
 * inputs collapse to the same float after exp/divide; which pairs collapse
 * depends on the libm behind ATen (Sleef on CPU, libdevice+MUFU on CUDA: the
 * reference's own CPU and GPU runs disagree there).  To make "bit-exact" well
 * defined on every input we pin ONE softmax arithmetic, implemented identically
 * (same IEEE-754 binary32 operations, in the same order, no contraction) here
 * and in the CUDA kernel:
 *     m   = max_c x_c
 *     d_c = x_c - m                               (binary32 subtract)
 *     e_c = spec_expf(d_c)                        (below)
 *     S   = ((e_0 + e_1) + e_2) + ...             (sequential binary32 adds)
 *     q_c = e_c / S                               (IEEE divide)
 *     pred = first c with q_c == max_c q_c ; if any d_c is NaN -> q is all-NaN -> pred = 0
 * spec_expf: Cody-Waite reduction + degree-5 Horner in explicit fmaf, exact 2^n scaling;
 * spec_expf(0) == 1 exactly; d < -87 -> 0.
 */
static inline float spec_expf(float d)
{
    if (!(d >= -87.0f)) return (d != d) ? d : 0.0f;
    float t = d * 1.44269504088896341f;
    float n = nearbyintf(t);
    float r = fmaf(n, -0.693359375f, d);
    r = fmaf(n, 2.12194440e-4f, r);
    float y = fmaf(1.9875691500e-4f, r, 1.3981999507e-3f);
    y = fmaf(y, r, 8.3334519073e-3f);
    y = fmaf(y, r, 4.1665795894e-2f);
    y = fmaf(y, r, 1.6666665459e-1f);
    y = fmaf(y, r, 5.0000001201e-1f);
    float z = r * r;
    y = fmaf(y, z, r);
    y = y + 1.0f;
    int32_t e = (int32_t)n + 127; /* n in [-126, 0] here */
    uint32_t bits = (uint32_t)e << 23;
    float scale;
    memcpy(&scale, &bits, 4);
    return y * scale;
}

float dcto_spec_expf(float d) { return spec_expf(d); }

static inline int spec_softmax_argmax(const float* xb, int C, int64_t HW)
{
    float m = xb[0];
    for (int c = 1; c < C; ++c) { float v = xb[c * HW]; if (v > m) m = v; }
    float S = 0.0f;
    int any_nan = 0;
    for (int c = 0; c < C; ++c) {
        float d = xb[c * HW] - m;
        if (d != d) any_nan = 1;
        float e = spec_expf(d);
        S = (c == 0) ? e : S + e;
    }
    if (any_nan) return 0;
    int best = 0;
    float qb = spec_expf(xb[0] - m) / S;
    for (int c = 1; c < C; ++c) {
        float q = spec_expf(xb[c * HW] - m) / S;
        if (q > qb) { qb = q; best = c; }
    }
    return best;
}

/* argmax of the RAW input, first index on ties, NaN counts as maximal (torch.max semantics):
 * IoU.add -> `_, predicted = predicted.max(1)`  generalframework/metrics/iou.py:64-65 */
static inline int raw_argmax(const float* xb, int C, int64_t HW)
{
    float m = xb[0];
    int best = 0;
    if (m != m) return 0;
    for (int c = 1; c < C; ++c) {
        float v = xb[c * HW];
        if (v != v) return c;
        if (v > m) { m = v; best = c; }
    }
    return best;
}

/* pred maps, for tests: mode 0 = spec softmax-argmax (Dice), 1 = raw argmax (IoU) */
void dcto_predict(const float* x, int64_t B, int C, int64_t HW, int mode, int64_t* pred)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t i = 0; i < HW; ++i) {
            const float* xb = x + b * C * HW + i;
            pred[b * HW + i] = mode == 0 ? spec_softmax_argmax(xb, C, HW) : raw_argmax(xb, C, HW);
        }
}

/*
 * DiceMeter.add counting -- generalframework/metrics/dice_meter.py:12-33,50-55 with
 * class2one_hot (utils/utils.py:187-198), intersection (:164-168):
 *   I[b,c] = #{pred==c & gt==c}, P[b,c] = #{pred==c}, G[b,c] = #{gt==c}
 * counts layout int64 [B][C][3] = (I, G, P).  Labels outside [0,C) make the
 * reference raise AssertionError (utils.py:190): they are counted in *bad and
 * excluded from G and I (P still counts the pixel).  '3d' Dice sums over b.
 */
void dcto_dice_counts(const float* x, const int64_t* gt, int64_t B, int C, int64_t HW,
                      int64_t* counts, int64_t* bad)
{
    int64_t nbad = 0;
    memset(counts, 0, sizeof(int64_t) * (size_t)(B * C * 3));
#pragma omp parallel for schedule(static) reduction(+ : nbad)
    for (int64_t b = 0; b < B; ++b) {
        int64_t* cb = counts + b * C * 3;
        for (int64_t i = 0; i < HW; ++i) {
            int p = spec_softmax_argmax(x + b * C * HW + i, C, HW);
            int64_t g = gt[b * HW + i];
            cb[p * 3 + 2] += 1;
            if (g < 0 || g >= C) { ++nbad; continue; }
            cb[g * 3 + 1] += 1;
            if (g == p) cb[p * 3 + 0] += 1;
        }
    }
    if (bad) *bad = nbad;
}

/* dice = (2*float32(I) + 1e-8) / (float32(G + P) + 1e-8) in float32 -- dice_meter.py:17-20 */
void dcto_dice_from_counts(const int64_t* counts, int64_t rows, int C, float* dice)
{
    for (int64_t r = 0; r < rows * C; ++r) {
        float inter = (float)counts[r * 3 + 0];
        float sum = (float)(counts[r * 3 + 1] + counts[r * 3 + 2]);
        dice[r] = (2.0f * inter + 1e-8f) / (sum + 1e-8f);
    }
}

/*
 * ConfusionMatrix.add -- generalframework/metrics/confusionmatrix.py:76-85 via IoU.add iou.py:43-69
 *   mask = (t >= 0) & (t < C); conf[t, pred] += 1 over masked pixels (rows = ground truth)
 * conf is int64 [C][C], ACCUMULATED into (the reference accumulates into np.int32).
 */
void dcto_confusion_from_scores(const float* x, const int64_t* gt, int64_t B, int C, int64_t HW, int64_t* conf)
{
    const int n = C * C;
#pragma omp parallel
    {
        int64_t local[n];
        memset(local, 0, sizeof(int64_t) * (size_t)n);
#pragma omp for collapse(2) schedule(static) nowait
        for (int64_t b = 0; b < B; ++b)
            for (int64_t i = 0; i < HW; ++i) {
                int64_t g = gt[b * HW + i];
                if (g < 0 || g >= C) continue;
                int p = raw_argmax(x + b * C * HW + i, C, HW);
                local[g * C + p] += 1;
            }
#pragma omp critical
        for (int j = 0; j < n; ++j) conf[j] += local[j];
    }
}

/* same, when IoU.add is handed an integer prediction map ([N,H,W] ints, iou.py:49-50).
 * numpy: x = predicted[mask] + C*target[mask]; bincount(x, minlength=C*C) then reshape(C,C)
 * (an out-of-range prediction would make bincount longer than C*C and the reference's
 *  `assert bincount_2d.size == C**2` fire; returned as the number of such pixels). */
int64_t dcto_confusion_from_labels(const int64_t* pred, const int64_t* gt, int64_t n, int C, int64_t* conf)
{
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t g = gt[i];
        if (g < 0 || g >= C) continue;
        int64_t key = pred[i] + (int64_t)C * g;
        if (key < 0 || key >= (int64_t)C * C) { ++bad; continue; }
        conf[key] += 1;
    }
    return bad;
}
