/* Development aid: counts rounds / evaluations of candidate parallel schedules for the region
 * voting stage, on real data dumped from the oracle.  Not part of the product or the tests. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef struct { uint8_t l, r, t, b; } arm_t;
static int W, H, D, N;
static arm_t* arms; static float *d_old, *d_new; static uint8_t* label;
static int ts = 20; static float th = 0.4f;
static int hist[512];
static long evals = 0;
static int last_peak, last_total;
static float vote(int p, const float* newv, const float* oldv) {
    int y = p / W, x = p % W;
    memset(hist, 0, sizeof(int) * D);
    arm_t a = arms[p];
    for (int t = -a.t; t <= a.b; t++) {
        int ri = (y + t) * W + x; arm_t a2 = arms[ri];
        for (int s = -a2.l; s <= a2.r; s++) {
            int before = (t < 0) || (t == 0 && s < 0);
            float d = before ? newv[ri + s] : oldv[ri + s];
            if (!isinf(d)) { long di = lroundf(d); if (di >= 0 && di < D) hist[di]++; }
        }
    }
    int best = 0, tot = 0, peak = 0;
    for (int d = 0; d < D; d++) { if (peak < hist[d]) { peak = hist[d]; best = d; } tot += hist[d]; }
    evals++; last_peak = peak; last_total = tot;
    if (peak > 0 && tot > ts && (float)peak / (float)tot > th) return (float)best;
    return INFINITY;
}
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); int hdr[3]; fread(hdr, 4, 3, f); W = hdr[0]; H = hdr[1]; D = hdr[2]; N = W * H;
    int mode = atoi(argv[2]); int G = argc > 3 ? atoi(argv[3]) : 32; int TILE = argc > 4 ? atoi(argv[4]) : 16;
    arms = malloc(N * 4); d_old = malloc(N * 4); d_new = malloc(N * 4); label = malloc(N);
    fread(arms, 4, N, f); fread(d_old, 4, N, f); fread(label, 1, N, f); fclose(f);
    memcpy(d_new, d_old, N * 4);
    int* lists[2]; int n[2] = {0, 0};
    for (int k = 0; k < 2; k++) { lists[k] = malloc(N * 4); for (int i = 0; i < N; i++) if (label[i] == k + 1) lists[k][n[k]++] = i; }
    int prune = argc > 5 ? atoi(argv[5]) : 0;
    if (prune) {
        uint8_t* hopeful = malloc(N); for (int i = 0; i < N; i++) hopeful[i] = label[i] != 0;
        long peel_evals = 0; int iters = 0, changed = 1;
        while (changed) { changed = 0; iters++;
            for (int i = 0; i < N; i++) if (hopeful[i]) { int y = i / W, x = i % W; arm_t a = arms[i]; int pot = 0; peel_evals++;
                for (int t = -a.t; t <= a.b; t++) { int ri = (y + t) * W + x; arm_t a2 = arms[ri]; for (int q = -a2.l; q <= a2.r; q++) pot += (prune == 1) ? 1 : ((!isinf(d_old[ri + q])) || hopeful[ri + q]); }
                if (pot <= ts) { hopeful[i] = 0; changed = 1; } }
            if (prune == 1) break; }
        int kept[2] = {0, 0};
        for (int k = 0; k < 2; k++) { int m = 0; for (int i = 0; i < n[k]; i++) if (hopeful[lists[k][i]]) lists[k][m++] = lists[k][i]; kept[k] = m; }
        printf("prune mode %d: iters %d peel_evals %ld kept %d/%d %d/%d\n", prune, iters, peel_evals, kept[0], n[0], kept[1], n[1]);
        n[0] = kept[0]; n[1] = kept[1]; }
    int tw = (W + TILE - 1) / TILE, thh = (H + TILE - 1) / TILE;
    int* stamp = calloc(tw * thh, 4); int* evalep = calloc(N, 4); int epoch = 1; int reach = 34;
    long rounds = 0; long checks = 0; long skipped_slack = 0;
    int* chg = calloc(tw * thh, 4); int* need = calloc(N, 4); int* snap = calloc(N, 4);
    #define BOXSUM(x,y) ({ int _s=0; for (int ty = (y - reach < 0 ? 0 : (y - reach) / TILE); ty <= (y + reach) / TILE && ty < thh; ty++) for (int tx = (x - reach < 0 ? 0 : (x - reach) / TILE); tx <= (x + reach) / TILE && tx < tw; tx++) _s += chg[ty*tw+tx]; _s; })

    for (int it = 0; it < 5; it++) for (int k = 0; k < 2; k++) {
        int* L = lists[k]; int cnt = n[k]; if (!cnt) continue;
        int sweep_rounds = 0; long e0 = evals;
        while (1) {
            int changed = 0;
            /* process in groups of G items: items of a group read the state as of the previous group (block Gauss-Seidel);
               G = cnt -> pure Jacobi; G = 1 -> sequential */
            for (int g0 = 0; g0 < cnt; g0 += G) {
                int g1 = g0 + G < cnt ? g0 + G : cnt;
                float res[4096]; int doit[4096];
                for (int i = g0; i < g1; i++) {
                    int p = L[i]; int y = p / W, x = p % W; checks++;
                    doit[i - g0] = (mode == 0) || stamp[(y / TILE) * tw + x / TILE] >= evalep[p];
                    if (doit[i - g0] && mode == 2 && isinf(d_new[p]) && need[p] > 0) {
                        int cur = BOXSUM(x, y);
                        if (cur - snap[p] < need[p]) { doit[i - g0] = 0; skipped_slack++; }
                    }
                    if (doit[i - g0]) { res[i - g0] = vote(p, d_new, d_old);
                        int nt = last_total <= ts ? ts + 1 - last_total : 0; float fr = th * last_total - last_peak; int nr = fr > 0 ? (int)floorf(fr) : 0;
                        need[p] = nt > nr ? nt : nr; if (need[p] < 1) need[p] = 1; snap[p] = BOXSUM(x, y); }
                }
                for (int i = g0; i < g1; i++) if (doit[i - g0]) {
                    int p = L[i]; int y = p / W, x = p % W; evalep[p] = epoch;
                    float r = res[i - g0];
                    if (memcmp(&r, &d_new[p], 4)) { d_new[p] = r; changed = 1; chg[(y / TILE) * tw + x / TILE]++;
                        for (int ty = (y - reach < 0 ? 0 : (y - reach) / TILE); ty <= (y + reach) / TILE && ty < thh; ty++)
                            for (int tx = (x - reach < 0 ? 0 : (x - reach) / TILE); tx <= (x + reach) / TILE && tx < tw; tx++) stamp[ty * tw + tx] = epoch; }
                }
            }
            rounds++; sweep_rounds++; epoch++;
            if (!changed) break;
        }
        int keep = 0;
        for (int i = 0; i < cnt; i++) { int p = L[i]; if (!isinf(d_new[p])) { d_old[p] = d_new[p]; int y = p / W, x = p % W; chg[(y / TILE) * tw + x / TILE]++;
                for (int ty = (y - reach < 0 ? 0 : (y - reach) / TILE); ty <= (y + reach) / TILE && ty < thh; ty++)
                    for (int tx = (x - reach < 0 ? 0 : (x - reach) / TILE); tx <= (x + reach) / TILE && tx < tw; tx++) stamp[ty * tw + tx] = epoch; }
            else L[keep++] = p; }
        epoch++;
        printf("it %d k %d: pending %d -> %d, rounds %d, evals %ld\n", it, k, cnt, keep, sweep_rounds, evals - e0);
        n[k] = keep;
    }
    printf("TOTAL rounds %ld evals %ld checks %ld slack-skips %ld\n", rounds, evals, checks, skipped_slack);
    FILE* o = fopen(argv[1], "ab"); fclose(o);
    /* checksum of result */
    unsigned long long cs = 0; for (int i = 0; i < N; i++) { uint32_t u; memcpy(&u, &d_old[i], 4); cs = cs * 1000003ull + u; }
    printf("checksum %llx\n", cs);
    return 0;
}
