// voting schedule simulator: counts evaluations / rounds / pixel visits under different dirty filters
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
static const char *sim_path(const char *name) { static char buf[4][512]; static int k; const char *d = getenv("ADC_SIM_DIR"); char *b = buf[k++ & 3]; snprintf(b, 512, "%s/%s", d ? d : "/tmp/sim", name); return b; }
#define W 450
#define H 375
#define N (W*H)
#define D 64
static float disp0[N], dref[N];
static uint8_t arms[N][4];
static uint16_t suph[N];
static int listv[2][N], nlist[2];
static int irv_ts = 20; static float irv_th = 0.4f;

static void *rd(const char *fn, void *dst, size_t bytes) { FILE *f = fopen(fn, "rb"); if (!f) { perror(fn); exit(1);} size_t n = fread(dst, 1, bytes, f); fclose(f); (void)n; return dst; }
static int rdlist(const char *fn, int *dst) { FILE *f = fopen(fn, "rb"); int xy[2]; int n = 0; while (fread(xy, 4, 2, f) == 2) dst[n++] = xy[1] * W + xy[0]; fclose(f); return n; }

// state: 255 invalid else value
static uint8_t st_old[N], st_new[N];
static long visits;
static int total_out, peak_out;
static int eval(int p, const uint8_t *o, const uint8_t *nw) {
    int y = p / W, x = p % W, hist[D] = {0};
    for (int t = -arms[p][2]; t <= arms[p][3]; t++) {
        int r = (y + t) * W + x;
        for (int s = -arms[r][0]; s <= arms[r][1]; s++) {
            int q = r + s;
            uint8_t v = (q < p) ? nw[q] : o[q];
            visits++;
            if (v != 255) hist[v]++;
        }
    }
    int best = 0, peak = 0, tot = 0;
    for (int b = 0; b < D; b++) { if (peak < hist[b]) { peak = hist[b]; best = b; } tot += hist[b]; }
    total_out = tot; peak_out = peak;
    if (tot > irv_ts && (float)peak / (float)tot > irv_th) return best;
    return 255;
}

int main(int argc, char **argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    int TILE = argc > 2 ? atoi(argv[2]) : 16; int BAND = argc > 3 ? atoi(argv[3]) : 100000; int fmode = argc > 4 ? atoi(argv[4]) : 0;
    rd(sim_path("disp.bin"), disp0, sizeof disp0); rd(sim_path("disp_vote.bin"), dref, sizeof dref);
    rd(sim_path("arms.bin"), arms, sizeof arms); rd(sim_path("suph.bin"), suph, sizeof suph);
    static int tmp[N];
    for (int k = 0; k < 2; k++) {
        int n = rdlist(k ? sim_path("oc.bin") : sim_path("mm.bin"), tmp);
        nlist[k] = 0;
        for (int i = 0; i < n; i++) if (suph[tmp[i]] > irv_ts) listv[k][nlist[k]++] = tmp[i];
    }
    for (int i = 0; i < N; i++) st_old[i] = isinf(disp0[i]) ? 255 : (uint8_t)lroundf(disp0[i]);
    memcpy(st_new, st_old, N);
    // sequential reference count
    if (mode == 9) {
        long ev = 0;
        for (int it = 0; it < 5; it++) for (int k = 0; k < 2; k++) {
            int m = 0;
            for (int i = 0; i < nlist[k]; i++) { int p = listv[k][i]; ev++; int r = eval(p, st_old, st_old); if (r != 255) st_old[p] = r; else listv[k][m++] = p; }
            nlist[k] = m;
        }
        int bad = 0; for (int i = 0; i < N; i++) { uint8_t e = isinf(dref[i]) ? 255 : (uint8_t)lroundf(dref[i]); bad += e != st_old[i]; }
        printf("sequential: evals %ld visits %ld mismatch_vs_ref %d\n", ev, visits, bad);
        return 0;
    }
    // fixed point with filters
    static int last_eval[N], chg_epoch[N], kmin[N], chgcount_snap[N];
    int tw = (W + TILE - 1) / TILE, th = (H + TILE - 1) / TILE;
    int *tstamp = calloc(tw * th, sizeof(int));
    int *tcount = calloc(tw * th, sizeof(int));   // cumulative number of changes stamped on the tile
    memset(last_eval, 0, sizeof last_eval); memset(chg_epoch, 0, sizeof chg_epoch);
    for (int i = 0; i < N; i++) { kmin[i] = 0; chgcount_snap[i] = 0; }
    int epoch = 1; long evals = 0, rounds = 0, skipped_k = 0, nonempty = 0, sumpar = 0;
    int reach = 34;
    for (int it = 0; it < 5; it++) for (int k = 0; k < 2; k++) {
        int n = nlist[k]; if (!n) continue;
int anyfill = 0;
if (mode == 4) {
        int i0 = 0;
        long maxpar = 0;
        while (i0 < n) {
          int i1 = i0; int yb = (listv[k][i0] / W) / BAND;
          while (i1 < n && (listv[k][i1] / W) / BAND == yb) i1++;
          while (1) {
            int changed = 0;
            static uint8_t snap[N]; memcpy(snap, st_new, N);
            static int chg_list[N]; int nchg = 0; long par = 0;
            for (int i = i0; i < i1; i++) {
                int p = listv[k][i], y = p / W, x = p % W;
                int dirty;
                if (fmode == 0) dirty = tstamp[(y / TILE) * tw + x / TILE] >= last_eval[p];
                else { dirty = last_eval[p] == 0;
                    for (int t = -arms[p][2]; t <= arms[p][3] && !dirty; t++) { int r = (y + t) * W + x; for (int s = -arms[r][0]; s <= arms[r][1]; s++) if (chg_epoch[r + s] >= last_eval[p]) { dirty = 1; break; } } }
                if (!dirty) continue;
                evals++; par++;
                int r = eval(p, st_old, snap);
                last_eval[p] = epoch;
                if (r != snap[p]) { st_new[p] = r; changed = 1; chg_list[nchg++] = p; }
            }
            for (int i = 0; i < nchg; i++) {
                int p = chg_list[i], y = p / W, x = p % W; chg_epoch[p] = epoch;
                int tx0 = (x - reach) / TILE, tx1 = (x + reach) / TILE, ty0 = y / TILE, ty1 = (y + reach) / TILE;
                if (x - reach < 0) tx0 = 0; if (tx1 >= tw) tx1 = tw - 1; if (ty1 >= th) ty1 = th - 1;
                for (int ty = ty0; ty <= ty1; ty++) for (int tx = tx0; tx <= tx1; tx++) { tstamp[ty * tw + tx] = epoch; tcount[ty * tw + tx]++; }
            }
            rounds++; epoch++; if (par) { nonempty++; sumpar += par; }
            if (!changed) break;
            anyfill = 1;
          }
          i0 = i1;
        }
} else {
        int anyfill = 0;
        while (1) {
            int changed = 0;
            // Jacobi-style round: evaluate using st_new as of round start for q<p?  GPU is asynchronous; emulate
            // "parallel round": all evals read a snapshot taken at round start.
            static uint8_t snap[N]; memcpy(snap, st_new, N);
            static int chg_list[N]; int nchg = 0;
            for (int i = 0; i < n; i++) {
                int p = listv[k][i], y = p / W, x = p % W;
                int dirty;
                if (mode == 0 || mode == 3) dirty = tstamp[(y / TILE) * tw + x / TILE] >= last_eval[p];
                else if (mode == 1) { // bbox exact
                    dirty = last_eval[p] == 0;
                    for (int t = -arms[p][2]; t <= arms[p][3] && !dirty; t++) { int r = (y + t) * W + x; for (int s = -34; s <= 34 && !dirty; s++) { if (x + s < 0 || x + s >= W) continue; if (chg_epoch[r + s] >= last_eval[p]) dirty = 1; } }
                } else { // exact region
                    dirty = last_eval[p] == 0;
                    for (int t = -arms[p][2]; t <= arms[p][3] && !dirty; t++) { int r = (y + t) * W + x; for (int s = -arms[r][0]; s <= arms[r][1]; s++) if (chg_epoch[r + s] >= last_eval[p]) { dirty = 1; break; } }
                }
                if (!dirty) continue;
                if (mode == 3 && last_eval[p] != 0 && snap[p] == 255) {
                    // count-of-changes bound: changes stamped on my tile since last eval
                    int c = tcount[(y / TILE) * tw + x / TILE] - chgcount_snap[p];
                    if (c < kmin[p]) { skipped_k++; continue; }
                }
                evals++;
                int r = eval(p, st_old, snap);
                last_eval[p] = epoch;
                if (mode == 3) {
                    chgcount_snap[p] = tcount[(y / TILE) * tw + x / TILE];
                    int km = 1;
                    if (r == 255) {
                        int a = irv_ts + 1 - total_out;                 // need total' >= ts+1
                        // ratio: (peak + c) / (total - c) > th  (most favourable)  -> c > (th*total - peak)/(1+th)
                        double bq = (irv_th * total_out - peak_out) / (1.0 + irv_th);
                        int b = (int)floor(bq - 1e-6);   // conservative: need c > bq  -> c >= floor(bq)+1; use floor(bq) to stay safe
                        if (b < 1) b = 1;
                        km = a > b ? a : b; if (km < 1) km = 1;
                    }
                    kmin[p] = km;
                }
                if (r != snap[p]) { st_new[p] = r; changed = 1; chg_list[nchg++] = p; }
            }
            for (int i = 0; i < nchg; i++) {
                int p = chg_list[i], y = p / W, x = p % W; chg_epoch[p] = epoch;
                int tx0 = (x - reach) / TILE, tx1 = (x + reach) / TILE, ty0 = y / TILE, ty1 = (y + reach) / TILE;
                if (x - reach < 0) tx0 = 0; if (tx1 >= tw) tx1 = tw - 1; if (ty1 >= th) ty1 = th - 1;
                for (int ty = ty0; ty <= ty1; ty++) for (int tx = tx0; tx <= tx1; tx++) { tstamp[ty * tw + tx] = epoch; tcount[ty * tw + tx]++; }
            }
            rounds++; epoch++;
            if (!changed) break;
            anyfill = 1;
        }
}
        if (!anyfill) continue;
        int m = 0;
        for (int i = 0; i < n; i++) {
            int p = listv[k][i], y = p / W, x = p % W;
            if (st_new[p] != 255) {
                st_old[p] = st_new[p]; chg_epoch[p] = epoch;
                int tx0 = (x - reach) / TILE, tx1 = (x + reach) / TILE, ty0 = (y - reach) / TILE, ty1 = (y + reach) / TILE;
                if (x - reach < 0) tx0 = 0; if (y - reach < 0) ty0 = 0; if (tx1 >= tw) tx1 = tw - 1; if (ty1 >= th) ty1 = th - 1;
                for (int ty = ty0; ty <= ty1; ty++) for (int tx = tx0; tx <= tx1; tx++) { tstamp[ty * tw + tx] = epoch; tcount[ty * tw + tx]++; }
            } else listv[k][m++] = p;
        }
        nlist[k] = m; epoch++;
    }
    int bad = 0; for (int i = 0; i < N; i++) { uint8_t e = isinf(dref[i]) ? 255 : (uint8_t)lroundf(dref[i]); bad += e != st_old[i]; }
    printf("band %d fmode %d nonempty_rounds %ld avg_par %.1f | ", BAND, fmode, nonempty, nonempty ? (double)sumpar / nonempty : 0.0); printf("mode %d tile %d: evals %ld rounds %ld visits %ld skipped_by_kmin %ld mismatch_vs_ref %d\n", mode, TILE, evals, rounds, visits, skipped_k, bad);
    return 0;
}
