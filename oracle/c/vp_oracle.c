/*
 * Plain-C restatement of the integer / index pieces of the SeisBench annotate path.
 * TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py ("PARITY UNPINNED").
 *
 * Follows:
 *   C.1  seisbench/models/base.py  WaveformModel._cut_fragments_array      -> vpo_window_starts
 *   C.4  seisbench/models/base.py  WaveformModel._reassemble_blocks_array  -> vpo_stack
 *        (NaN slot buffer, slot = i % coverage, np.nanmean / np.nanmax; the fp32 sum follows
 *         NumPy's pairwise_sum order so that results are bit-identical to np.nanmean)
 *   C.6  obspy/signal/trigger.py   trigger_onset + argmax, as restated in the reference at
 *        /root/reference/volpick/model/eval_taks0.py:46-56                 -> vpo_picks
 *
 * Build:  gcc -O2 -fPIC -shared -o oracle/_build/libvp_oracle.so oracle/c/vp_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* C.1: starts = arange(0, N-L+1, L-overlap); append N-L if the last window stops short of N. */
int64_t vpo_window_starts(int64_t n, int64_t len, int64_t overlap, int64_t *out, int64_t cap) {
    int64_t stride = len - overlap, cnt = 0, s;
    if (stride <= 0 || n < len) return 0;
    for (s = 0; s <= n - len; s += stride) {
        if (out && cnt < cap) out[cnt] = s;
        cnt++;
    }
    s -= stride;
    if (s + len < n) {
        if (out && cnt < cap) out[cnt] = n - len;
        cnt++;
    }
    return cnt;
}

/* NumPy's float pairwise sum (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum). */
static float np_pairwise_sum_f32(const float *a, int64_t n) {
    if (n < 8) {
        float res = -0.0f; /* NumPy starts from -0.0 so that sum([-0.0]) == -0.0 */
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8], res;
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
    }
}

/*
 * C.4.  y: (B, C, L) planar window predictions (NOT yet blinded); starts sorted ascending.
 * out: (C, pred_len) planar, pred_len = max(starts) + L.  mode 0 = avg (nanmean), 1 = max (nanmax).
 * Samples inside the blinding margins [0,b0) and [L-b1,L) of a window count as NaN.
 */
int vpo_stack(const float *y, const int64_t *starts, int64_t nwin, int64_t len, int nch, int64_t cov,
              int64_t b0, int64_t b1, int mode, float *out, int64_t pred_len) {
    float *slot = (float *)malloc(sizeof(float) * (size_t)cov);
    float *vals = (float *)malloc(sizeof(float) * (size_t)cov);
    if (!slot || !vals) return -1;
    for (int c = 0; c < nch; c++) {
        int64_t lo = 0; /* first window that can still cover n */
        for (int64_t n = 0; n < pred_len; n++) {
            for (int64_t k = 0; k < cov; k++) slot[k] = NAN;
            while (lo < nwin && starts[lo] + len <= n) lo++;
            for (int64_t i = lo; i < nwin && starts[i] <= n; i++) {
                int64_t off = n - starts[i];
                if (off >= len) continue;
                /* later windows overwrite the slot, exactly like the buffer assignment */
                slot[i % cov] = (off < b0 || off >= len - b1) ? NAN : y[(i * nch + c) * len + off];
            }
            int64_t cnt = 0;
            float res;
            if (mode == 0) {
                for (int64_t k = 0; k < cov; k++) {
                    if (isnan(slot[k])) vals[k] = 0.0f; else { vals[k] = slot[k]; cnt++; }
                }
                float tot = 0.0f + np_pairwise_sum_f32(vals, cov);
                res = cnt ? (float)((double)tot / (double)cnt) : NAN;
            } else {
                res = NAN;
                for (int64_t k = 0; k < cov; k++)
                    if (!isnan(slot[k]) && (isnan(res) || slot[k] > res)) res = slot[k];
            }
            out[(int64_t)c * pred_len + n] = res;
        }
    }
    free(slot);
    free(vals);
    return 0;
}

/*
 * C.6 run form: one trigger per maximal run of x > thr_off containing a sample > thr_on.
 * Returns the number of triggers found (may exceed cap; only cap are written).
 */
int64_t vpo_picks(const float *x, int64_t n, float thr_on, float thr_off, int64_t *s0, int64_t *s1,
                  int64_t *speak, float *vpeak, int64_t cap) {
    int64_t cnt = 0, i = 0;
    while (i < n) {
        if (x[i] > thr_off) {
            int64_t j = i, on = -1, pk = -1;
            float best = 0.0f;
            while (j < n && x[j] > thr_off) {
                if (on < 0 && x[j] > thr_on) { on = j; pk = j; best = x[j]; }
                else if (on >= 0 && x[j] > best) { pk = j; best = x[j]; } /* strict > : first maximum */
                j++;
            }
            if (on >= 0) {
                if (cnt < cap) { s0[cnt] = on; s1[cnt] = j - 1; speak[cnt] = pk; vpeak[cnt] = best; }
                cnt++;
            }
            i = j;
        } else {
            i++;
        }
    }
    return cnt;
}
