// push-based voting simulation: counts changes, inverse-region checks, pushes, re-derives
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
static void *rd(const char *fn, void *dst, size_t bytes) { FILE *f = fopen(fn, "rb"); size_t n = fread(dst, 1, bytes, f); fclose(f); (void)n; return dst; }
static int rdlist(const char *fn, int *dst) { FILE *f = fopen(fn, "rb"); int xy[2]; int n = 0; while (fread(xy, 4, 2, f) == 2) dst[n++] = xy[1] * W + xy[0]; fclose(f); return n; }
static uint8_t st_old[N], st_new[N];
static int slot[N]; static int8_t which[N];   // list id of pending pixel or -1
static uint16_t (*hist)[D];
static int ihl[N], ihr[N], ivt[N], ivb[N];   // inverse extents
static uint8_t dirty[N];
static long checks, pushes, derives, changes, visits;

static void push_all(int q, int a, int b, int cur_k, int phase) {
    // phase 0: during rounds of list cur_k: push to other-list pixels and to current-list p > q
    // phase 1: commit: push (invalid -> b) to current-list p < q   (a == 255)
    int qy = q / W, qx = q % W;
    for (int px = ihl[q]; px <= ihr[q]; px++) {
        int r = qy * W + px; checks++;
        if (!(px - arms[r][0] <= qx && qx <= px + arms[r][1])) continue;
        for (int py = ivt[r]; py <= ivb[r]; py++) {
            int p = py * W + px; checks++;
            if (which[p] < 0) continue;
            if (!(py - arms[p][2] <= qy && qy <= py + arms[p][3])) continue;
            int now;
            if (which[p] != cur_k) now = (phase == 1);
            else now = (phase == 0) ? (p > q) : (p < q);
            if (!now) continue;
            if (a != 255) hist[slot[p]][a]--;
            if (b != 255) hist[slot[p]][b]++;
            pushes++; dirty[p] = 1;
        }
    }
}
int main() {
    rd(sim_path("disp.bin"), disp0, sizeof disp0); rd(sim_path("disp_vote.bin"), dref, sizeof dref);
    rd(sim_path("arms.bin"), arms, sizeof arms); rd(sim_path("suph.bin"), suph, sizeof suph);
    static int tmp[N]; int ns = 0;
    memset(which, -1, sizeof which);
    for (int k = 0; k < 2; k++) {
        int n = rdlist(k ? sim_path("oc.bin") : sim_path("mm.bin"), tmp);
        nlist[k] = 0;
        for (int i = 0; i < n; i++) if (suph[tmp[i]] > irv_ts) { listv[k][nlist[k]++] = tmp[i]; slot[tmp[i]] = ns++; which[tmp[i]] = k; }
    }
    hist = calloc(ns, sizeof *hist);
    for (int i = 0; i < N; i++) st_old[i] = isinf(disp0[i]) ? 255 : (uint8_t)lroundf(disp0[i]);
    memcpy(st_new, st_old, N);
    // inverse extents
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) { int q = y * W + x; ihl[q] = x; ihr[q] = x; ivt[q] = y; ivb[q] = y; }
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        int p = y * W + x;
        for (int s = -arms[p][0]; s <= arms[p][1]; s++) { int q = p + s; if (x < ihl[q]) ihl[q] = x; if (x > ihr[q]) ihr[q] = x; }
        for (int t = -arms[p][2]; t <= arms[p][3]; t++) { int q = p + t * W; if (y < ivt[q]) ivt[q] = y; if (y > ivb[q]) ivb[q] = y; }
    }
    // note: ivt/ivb defined per pixel r=(px,qy): rows py whose v-arm covers qy
    // initial hists
    for (int k = 0; k < 2; k++) for (int i = 0; i < nlist[k]; i++) {
        int p = listv[k][i], y = p / W, x = p % W;
        for (int t = -arms[p][2]; t <= arms[p][3]; t++) { int r = (y + t) * W + x; for (int s = -arms[r][0]; s <= arms[r][1]; s++) { uint8_t v = st_old[r + s]; visits++; if (v != 255) hist[slot[p]][v]++; } }
        dirty[p] = 1;
    }
    long rounds = 0;
    for (int it = 0; it < 5; it++) for (int k = 0; k < 2; k++) {
        int n = nlist[k]; if (!n) continue;
        int anyfill = 0;
        while (1) {
            static int cq[N], ca[N], cb[N]; int nc = 0;
            for (int i = 0; i < n; i++) {
                int p = listv[k][i]; if (!dirty[p]) continue;
                dirty[p] = 0; derives++;
                int best = 0, peak = 0, tot = 0; uint16_t *h = hist[slot[p]];
                for (int b = 0; b < D; b++) { if (peak < h[b]) { peak = h[b]; best = b; } tot += h[b]; }
                int r = (tot > irv_ts && (float)peak / (float)tot > irv_th) ? best : 255;
                if (r != st_new[p]) { cq[nc] = p; ca[nc] = st_new[p]; cb[nc] = r; nc++; }
            }
            rounds++;
            if (!nc) break;
            anyfill = 1; changes += nc;
            for (int i = 0; i < nc; i++) { st_new[cq[i]] = cb[i]; push_all(cq[i], ca[i], cb[i], k, 0); }
        }
        if (!anyfill) continue;
        int m = 0;
        for (int i = 0; i < n; i++) {
            int p = listv[k][i];
            if (st_new[p] != 255) { st_old[p] = st_new[p]; which[p] = -1; changes++; }
        }
        for (int i = 0; i < n; i++) { int p = listv[k][i]; if (st_new[p] != 255) push_all(p, 255, st_new[p], k, 1); else listv[k][m++] = p; }
        nlist[k] = m;
    }
    int bad = 0; for (int i = 0; i < N; i++) { uint8_t e = isinf(dref[i]) ? 255 : (uint8_t)lroundf(dref[i]); bad += e != st_old[i]; }
    printf("push: slots %d init_visits %ld rounds %ld derives %ld changes(incl commits) %ld checks %ld pushes %ld mismatch_vs_ref %d\n", ns, visits, rounds, derives, changes, checks, pushes, bad);
    return 0;
}
