// CPU model of K1 query strategies: counts dependent shared-memory steps per query and the
// max over 32 consecutive positions (a warp pass).  Experiment tool, not product.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../../lzs-compression_b200/csrc/corpus.h"

static const int W = 2047;
static int SLOTS_LOG = 11;
static int TAGBITS = 5;

static int HASHMODE = 0;
static int GALLOP = 2;
static int LEVELS = 0x1FFC;   /* bit k set: level k has a table */
static inline uint32_t rolling_hash(const uint8_t *p, int k)
{
    // groups of levels as the build warps compute them: {2,3,4} {5,6,7} {8,9,10} {11,12}
    const uint32_t C = 0x9E3779B1u, C2 = 0x85EBCA77u;
    uint32_t w0 = p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24;
    uint32_t w1 = p[4] | p[5] << 8 | p[6] << 16 | (uint32_t)p[7] << 24;
    uint32_t w2 = p[8] | p[9] << 8 | p[10] << 16 | (uint32_t)p[11] << 24;
    uint32_t h;
    int from;
    if (k <= 4) { h = (w0 & 0xFFFFu) * C; from = 2; }
    else if (k <= 7) { h = w0 * C; from = 4; }
    else if (k <= 10) { h = ((w0 * C) ^ (w1 & 0xFFFFFFu)) * C2; from = 7; }
    else { h = ((((w0 * C) ^ w1) * C2) ^ (w2 & 0xFFFFu)) * C; from = 10; }
    for (int b = from; b < k; b++) {
        uint32_t w = b < 4 ? w0 : b < 8 ? w1 : w2;
        h = (h ^ (w & (0xFFu << (8 * (b & 3))))) * C;
    }
    return h;
}
static inline uint32_t gram_hash(const uint8_t *p, int k)
{
    if (HASHMODE) return rolling_hash(p, k);
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    for (int b = 0; b < k; b++) {
        if (b < 4) w0 |= (uint32_t)p[b] << (8 * b);
        else if (b < 8) w1 |= (uint32_t)p[b] << (8 * (b - 4));
        else w2 |= (uint32_t)p[b] << (8 * (b - 8));
    }
    uint32_t h = (w0 * 0x9E3779B1u) ^ (w1 * 0x85EBCA77u) ^ (w2 * 0xC2B2AE3Du);
    h ^= h >> 15;
    h *= 0x27D4EB2Fu;
    return h;
}
static inline int lcp(const uint8_t *a, const uint8_t *b, int M)
{
    int l = 0;
    while (l < M && a[l] == b[l]) l++;
    return l;
}

struct Level {
    std::vector<int32_t> head;          // slot -> last position (-1 none)
    std::vector<int32_t> d1;            // pos -> distance to previous in slot (0 none / > W)
    std::vector<int32_t> skip;          // pos -> distance to nearest previous slot entry with different tag (0 none)
    std::vector<uint32_t> tag;
};

struct Stats {
    double q = 0, steps = 0, maxsum = 0, passes = 0, foreign = 0, verify = 0, resolvedA = 0;
    double walkers = 0;   // queries that need any chain hop beyond phase A
    double maxB = 0, passesB = 0, stepsB = 0;
};

int main(int argc, char **argv)
{
    int kind = argc > 1 ? atoi(argv[1]) : 0;
    int nstreams = argc > 2 ? atoi(argv[2]) : 4;
    SLOTS_LOG = argc > 3 ? atoi(argv[3]) : 11;
    TAGBITS = argc > 4 ? atoi(argv[4]) : 5;
    HASHMODE = argc > 5 ? atoi(argv[5]) : 0;
    GALLOP = argc > 6 ? atoi(argv[6]) : 2;
    LEVELS = argc > 7 ? (int)strtol(argv[7], 0, 0) : 0x1FFC;
    const int n = 65536;
    std::vector<uint8_t> buf(n + 64, 0);
    Stats cur, v1, v2, v3, v4, v5;
    double lenhist[13] = {0};
    for (int s = 0; s < nstreams; s++) {
        memset(buf.data(), 0, buf.size());
        lzs_corpus_fill(buf.data(), n, 0x5EED0002ull, (uint64_t)s * 3 + (kind == 3 ? 0 : 0), kind);
        std::vector<Level> L(13);
        const uint32_t slots = 1u << SLOTS_LOG;
        for (int k = 2; k <= 12; k++) {
            L[k].head.assign(slots, -1);
            L[k].d1.assign(n, 0);
            L[k].skip.assign(n, 0);
            L[k].tag.assign(n, 0);
        }
        // build all
        for (int i = 0; i < n; i++)
            for (int k = 2; k <= 12; k++) {
                uint32_t h = gram_hash(&buf[i], k);
                uint32_t slot = h >> (32 - SLOTS_LOG);
                uint32_t tag = (h >> (HASHMODE ? (32 - SLOTS_LOG - TAGBITS) : 10)) & ((1u << TAGBITS) - 1);
                int32_t q = L[k].head[slot];
                L[k].head[slot] = i;
                L[k].tag[i] = tag;
                int d = (q >= 0 && i - q <= W) ? i - q : 0;
                L[k].d1[i] = d;
                if (d == 0) L[k].skip[i] = 0;
                else if (L[k].tag[q] != tag) L[k].skip[i] = d;
                else {
                    int sk = L[k].skip[q];
                    L[k].skip[i] = (sk && d + sk <= W) ? d + sk : 0;
                }
            }
        // queries
        std::vector<int> c0(n), c1(n), c2(n), c3(n), c4(n), c5(n), needB(n), lenbest(n);
        for (int i = 1; i + 12 <= n; i++) {
            const int M = 12, maxd = std::min(W, i);
            // ---- V0: current upward walk; steps = chain hops (LDS of entries)
            {
                int k = 2, steps = 0, best = 0;
                for (;;) {
                    int tot = 0, d = L[k].d1[i];
                    bool found = false;
                    steps++;                       // own entry read
                    while (d && tot + d <= maxd) {
                        tot += d;
                        int j = i - tot;
                        steps++;                   // entry at j
                        d = L[k].d1[j];
                        if (L[k].tag[j] != L[k].tag[i]) { cur.foreign++; continue; }
                        int l = lcp(&buf[i], &buf[j], M);
                        cur.verify++;
                        if (l < k) continue;
                        best = l; found = true; break;
                    }
                    if (!found || best >= M) break;
                    k = best + 1;
                }
                c0[i] = steps; lenbest[i] = best;
                lenhist[best]++;
            }
            // ---- V1: E-mask (pred tag known from head), start at top level with immediate tag hit
            // ---- V2: same + collapse skip on foreign entries
            for (int variant = 1; variant <= 2; variant++) {
                int steps = 0, hops = 0;
                // phase A: top E level
                int kE = 0;
                for (int k = 12; k >= 2; k--) {
                    int d = L[k].d1[i];
                    if (d && d <= maxd && L[k].tag[i - d] == L[k].tag[i]) { kE = k; break; }
                }
                steps++;                           // mask read
                int best = 0, k = 2;
                bool resolved = true;
                if (kE) {
                    int d = L[kE].d1[i];
                    steps += 2;                    // link read + verify
                    int l = lcp(&buf[i], &buf[i - d], M);
                    if (l >= kE) { best = l; k = l + 1; }
                    else { k = 2; }                // rare: tag collision; fall back to the general walk from 2
                }
                // general upward walk from level k
                while (best < M) {
                    int tot = 0, d = L[k].d1[i];
                    bool found = false;
                    // immediate-pred knowledge from the mask: level empty -> no step at all
                    if (!(d && d <= maxd)) break;
                    resolved = false;
                    steps++;                       // own entry read
                    uint32_t mytag = L[k].tag[i];
                    while (d && tot + d <= maxd) {
                        tot += d;
                        int j = i - tot;
                        steps++; hops++;
                        if (L[k].tag[j] != mytag) {
                            d = (variant == 2) ? L[k].skip[j] : L[k].d1[j];
                            continue;
                        }
                        d = L[k].d1[j];
                        int l = lcp(&buf[i], &buf[j], M);
                        if (l < k) continue;
                        best = l; found = true; break;
                    }
                    if (!found) break;
                    k = best + 1;
                }
                if (variant == 1) { c1[i] = steps; if (resolved) v1.resolvedA++; }
                else { c2[i] = steps; if (resolved) v2.resolvedA++; needB[i] = resolved ? 0 : (steps - (kE ? 3 : 1)); }
            }
            // ---- V4: upward for the first GALLOP probes, then binary search over the remaining levels
            {
                int steps = 0, best = 0, probes = 0;
                int lo = 2, hi = M;                 // levels still undecided: lo..hi (hi < lo: done)
                while (lo <= hi && best < M) {
                    int k = (probes < GALLOP) ? lo : (lo + hi + 1) / 2;
                    probes++;
                    int tot = 0, d = L[k].d1[i];
                    bool found = false;
                    steps++;
                    while (d && tot + d <= maxd) {
                        tot += d;
                        int j = i - tot;
                        steps++;
                        d = L[k].d1[j];
                        if (L[k].tag[j] != L[k].tag[i]) continue;
                        int l = lcp(&buf[i], &buf[j], M);
                        if (l < k) continue;
                        best = l; found = true; break;
                    }
                    if (found) lo = best + 1; else hi = k - 1;
                }
                c4[i] = steps;
                if (best != (int)lenbest[i]) { fprintf(stderr, "V4 mismatch at %d: %d vs %d\n", i, best, lenbest[i]); exit(1); }
            }
            // ---- V5: sparse levels: level k is looked up on the chain of the largest BUILT level <= k
            {
                int k = 2, steps = 0, best = 0;
                for (;;) {
                    int kb = k;
                    while (!(LEVELS >> kb & 1)) kb--;
                    int tot = 0, d = L[kb].d1[i];
                    bool found = false;
                    steps++;
                    while (d && tot + d <= maxd) {
                        tot += d;
                        int j = i - tot;
                        steps++;
                        d = L[kb].d1[j];
                        if (L[kb].tag[j] != L[kb].tag[i]) continue;
                        int l = lcp(&buf[i], &buf[j], M);
                        if (l < k) continue;
                        best = l; found = true; break;
                    }
                    if (!found || best >= M) break;
                    k = best + 1;
                }
                c5[i] = steps;
                if (best != (int)lenbest[i]) { fprintf(stderr, "V5 mismatch at %d: %d vs %d\n", i, best, lenbest[i]); exit(1); }
            }
            // ---- V3: plain upward walk (V0) but with collapse skip
            {
                int k = 2, steps = 0, best = 0;
                for (;;) {
                    int tot = 0, d = L[k].d1[i];
                    bool found = false;
                    steps++;
                    while (d && tot + d <= maxd) {
                        tot += d;
                        int j = i - tot;
                        steps++;
                        if (L[k].tag[j] != L[k].tag[i]) { d = L[k].skip[j]; continue; }
                        d = L[k].d1[j];
                        int l = lcp(&buf[i], &buf[j], M);
                        if (l < k) continue;
                        best = l; found = true; break;
                    }
                    if (!found || best >= M) break;
                    k = best + 1;
                }
                c3[i] = steps;
            }
        }
        auto acc = [&](Stats &st, std::vector<int> &c) {
            for (int b = 32; b + 32 <= n - 12; b += 32) {
                int mx = 0;
                for (int l = 0; l < 32; l++) { st.steps += c[b + l]; st.q++; mx = std::max(mx, c[b + l]); }
                st.maxsum += mx; st.passes++;
            }
        };
        acc(cur, c0); acc(v1, c1); acc(v2, c2); acc(v3, c3); acc(v4, c4); acc(v5, c5);
        // phase B for V2: compact unresolved positions of a 448 tile into groups of 32
        for (int t = 32; t + 448 <= n - 12; t += 448) {
            std::vector<int> list;
            for (int l = 0; l < 448; l++) if (needB[t + l] > 0) list.push_back(needB[t + l]);
            v2.walkers += list.size();
            for (size_t g = 0; g < list.size(); g += 32) {
                int mx = 0;
                for (size_t l = g; l < std::min(list.size(), g + 32); l++) { mx = std::max(mx, list[l]); v2.stepsB += list[l]; }
                v2.maxB += mx; v2.passesB++;
            }
        }
    }
    const char *names[] = {"text", "binary", "random", "mixed", "packet"};
    printf("kind=%s slots=2^%d tag=%d\n", names[kind], SLOTS_LOG, TAGBITS);
    auto pr = [&](const char *nm, Stats &s) {
        printf("  %-28s steps/query %.2f  warp-max steps %.2f  (per pos %.3f)  foreign/q %.2f  resolvedA %.1f%%\n", nm,
               s.steps / s.q, s.maxsum / s.passes, s.maxsum / s.passes / 32, s.foreign / s.q, 100 * s.resolvedA / s.q);
    };
    pr("V0 current upward", cur);
    pr("V3 upward+collapse", v3);
    pr("V4 gallop+binary", v4);
    { char nm[64]; snprintf(nm, sizeof nm, "V5 sparse levels %03x", LEVELS >> 2); pr(nm, v5); }
    pr("V1 mask start", v1);
    pr("V2 mask start+collapse", v2);
    printf("  V2 two-phase: walkers %.1f%%  B steps/walker %.2f  B warp-max %.2f  B passes per 448-tile %.2f -> B max-steps per pos %.3f\n",
           100 * v2.walkers / v2.q, v2.stepsB / std::max(1.0, v2.walkers), v2.maxB / std::max(1.0, v2.passesB),
           v2.passesB / (v2.q / 448), v2.maxB / v2.q);
    printf("  best-length histogram:");
    double tot = 0; for (int l = 0; l <= 12; l++) tot += lenhist[l];
    for (int l = 0; l <= 12; l++) printf(" %d:%.1f%%", l, 100 * lenhist[l] / tot);
    printf("\n");
    return 0;
}
