/*
 * ddz_oracle.c -- CPU restatement of the reference's Doudizhu env hot path.
 * TEST INFRASTRUCTURE ONLY -- see ddz_oracle.h for who may use it and for the
 * parity status ("parity unpinned" for the absent natives; Python layer and
 * rules pinned by tests/golden/).
 *
 * Every function cites the reference file:line (relative to the reference
 * root) that it restates.  Nothing here is copied from the reference: the
 * reference has no C source at all; the rules are re-expressed from its
 * Python (rule_based/utils/card.py, envi.py, server/core.py, game.py).
 */
#define _POSIX_C_SOURCE 200809L
#include "ddz_oracle.h"
#include <pthread.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* universe of moves                                                    */
/* ------------------------------------------------------------------ */
typedef struct {
    int8_t cnt[15];
    int8_t cat, len, val, extra;
    uint64_t packed;
} act_t;

static act_t U[DDZ_REF_NUNIVERSE + 8];
static int UN = 0;
#define HSIZE 65536
static int32_t htab[HSIZE];
static pthread_once_t once = PTHREAD_ONCE_INIT;

uint64_t ddz_ref_pack(const int8_t c[15]) {
    uint64_t p = 0;
    for (int i = 0; i < 15; i++) p |= (uint64_t)(c[i] & 15) << (4 * i);
    return p;
}
void ddz_ref_unpack(uint64_t p, int8_t c[15]) {
    for (int i = 0; i < 15; i++) c[i] = (int8_t)((p >> (4 * i)) & 15);
}
static uint32_t hash64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return (uint32_t)x & (HSIZE - 1);
}
static void push(const int8_t c[15], int cat, int len, int val, int extra) {
    act_t* a = &U[UN];
    memcpy(a->cnt, c, 15);
    a->cat = (int8_t)cat; a->len = (int8_t)len; a->val = (int8_t)val; a->extra = (int8_t)extra;
    a->packed = ddz_ref_pack(c);
    uint32_t h = hash64(a->packed);
    while (htab[h] >= 0) h = (h + 1) & (HSIZE - 1);
    htab[h] = UN++;
}

/* lexicographic k-combinations of pool[0..n), the order itertools.combinations yields
 * (card.py:115,128,141,151) */
typedef void (*combo_cb)(const int* pick, int k, void* ctx);
static void combos(const int* pool, int n, int k, combo_cb cb, void* ctx) {
    int idx[8], pick[8] = {0};
    if (k > n) return;
    for (int i = 0; i < k; i++) idx[i] = i;
    for (;;) {
        for (int i = 0; i < k; i++) pick[i] = pool[idx[i]];
        cb(pick, k, ctx);
        int i = k - 1;
        while (i >= 0 && idx[i] == n - k + i) i--;
        if (i < 0) return;
        idx[i]++;
        for (int j = i + 1; j < k; j++) idx[j] = idx[j - 1] + 1;
    }
}

typedef struct { int8_t base[15]; int cat, len, val, kmult, filter_rocket; } gen_ctx;
static void gen_cb(const int* pick, int k, void* vctx) {
    gen_ctx* g = (gen_ctx*)vctx;
    int8_t c[15];
    memcpy(c, g->base, 15);
    int jokers = 0;
    for (int i = 0; i < k; i++) { c[pick[i]] += g->kmult; if (pick[i] >= 13) jokers++; }
    /* card.py:116 and :142 drop the combination {*,$} when it is exactly two kickers;
     * r.get_moves does emit those (server/mcts/get_moves.py:22-34) -> universe keeps them, flagged */
    int extra = (g->filter_rocket && k == 2 && jokers == 2);
    push(c, g->cat, g->len, g->val, extra);
}

/* card.py:34-159 get_action_space(), same order; value/len as CardGroup.analyze assigns them
 * (card.py:372-527: value = lowest main rank index, len = sequence length, 1 for non-lines) */
static void build_universe(void) {
    int8_t c[15];
    for (int i = 0; i < HSIZE; i++) htab[i] = -1;
    UN = 0;
    memset(c, 0, 15);
    push(c, 0, 1, 0, 0);                                   /* :35 pass */
    for (int r = 0; r < 15; r++) { memset(c, 0, 15); c[r] = 1; push(c, 1, 1, r, 0); }   /* :41 */
    for (int m = 2; m <= 4; m++)                           /* :47 :54 :61 pair / triple / bomb */
        for (int r = 0; r < 13; r++) { memset(c, 0, 15); c[r] = (int8_t)m; push(c, m, 1, r, 0); }
    for (int r = 0; r < 13; r++)                           /* :68 3+1 */
        for (int e = 0; e < 15; e++) if (e != r) { memset(c, 0, 15); c[r] = 3; c[e] = 1; push(c, 5, 1, r, 0); }
    for (int r = 0; r < 13; r++)                           /* :77 3+2 */
        for (int e = 0; e < 13; e++) if (e != r) { memset(c, 0, 15); c[r] = 3; c[e] = 2; push(c, 6, 1, r, 0); }
    /* :86 :94 :102 single / double / triple sequences: start 3..A, end exclusive <= 12 */
    static const int lmin[3] = {5, 3, 2}, lmax[3] = {12, 10, 6};
    for (int t = 0; t < 3; t++)
        for (int s = 0; s < 12; s++)
            for (int L = lmin[t]; L <= lmax[t] && s + L <= 12; L++) {
                memset(c, 0, 15);
                for (int i = s; i < s + L; i++) c[i] = (int8_t)(t + 1);
                push(c, 7 + t, L, s, 0);
            }
    /* :110 3+1 sequences (len 2..5), :122 3+2 sequences (len 2..4) */
    for (int t = 0; t < 2; t++)
        for (int s = 0; s < 12; s++)
            for (int L = 2; L <= (t == 0 ? 5 : 4) && s + L <= 12; L++) {
                gen_ctx g; memset(g.base, 0, 15);
                for (int i = s; i < s + L; i++) g.base[i] = 3;
                int pool[15], n = 0;
                for (int r = 0; r < (t == 0 ? 15 : 13); r++) if (r < s || r >= s + L) pool[n++] = r;
                g.cat = 10 + t; g.len = L; g.val = s; g.kmult = t + 1; g.filter_rocket = (t == 0);
                combos(pool, n, L, gen_cb, &g);
            }
    memset(c, 0, 15); c[13] = 1; c[14] = 1; push(c, 12, 1, 100, 0);  /* :133 rocket; value 100 card.py:383 */
    /* :138 4+1+1, :148 4+2+2 */
    for (int t = 0; t < 2; t++)
        for (int r = 0; r < 13; r++) {
            gen_ctx g; memset(g.base, 0, 15); g.base[r] = 4;
            int pool[15], n = 0;
            for (int e = 0; e < (t == 0 ? 15 : 13); e++) if (e != r) pool[n++] = e;
            g.cat = 13 + t; g.len = 1; g.val = r; g.kmult = t + 1; g.filter_rocket = (t == 0);
            combos(pool, n, 2, gen_cb, &g);
        }
}
static void ensure(void) { pthread_once(&once, build_universe); }

int ddz_ref_universe_size(void) { ensure(); return UN; }
int ddz_ref_universe_get(int i, int8_t counts[15], int* cat, int* len, int* val, int* is_extra) {
    ensure();
    if (i < 0 || i >= UN) return -1;
    memcpy(counts, U[i].cnt, 15);
    if (cat) *cat = U[i].cat;
    if (len) *len = U[i].len;
    if (val) *val = U[i].val;
    if (is_extra) *is_extra = U[i].extra;
    return 0;
}
static int lookup(uint64_t packed) {
    uint32_t h = hash64(packed);
    while (htab[h] >= 0) {
        if (U[htab[h]].packed == packed) return htab[h];
        h = (h + 1) & (HSIZE - 1);
    }
    return -1;
}
int ddz_ref_classify(const int8_t counts[15], int* cat, int* len, int* val) {
    ensure();
    for (int i = 0; i < 15; i++) if (counts[i] < 0 || counts[i] > 4) return -1;
    int i = lookup(ddz_ref_pack(counts));
    if (i < 0) return -1;
    if (cat) *cat = U[i].cat;
    if (len) *len = U[i].len;
    if (val) *val = U[i].val;
    return i;
}

/* card.py:307-325 CardGroup.bigger_than, branch for branch */
int ddz_ref_bigger_than(int cat, int len, int val, int gcat, int glen, int gval) {
    if (cat == 0) return gcat != 0;
    if (gcat == 0) return 1;
    if (gcat == 12) return 0;
    if (cat == 12) return 1;
    if (gcat == 4) return (cat == 4 && val > gval);
    if (cat == 4 || (cat == gcat && len == glen && val > gval)) return 1;
    return 0;
}

static int is_zero15(const int8_t* c) { for (int i = 0; i < 15; i++) if (c[i]) return 0; return 1; }

/* legal set = containment AND beats (server/rule_utils/utils.py:21-38; rule_based/utils/utils.py:16-22);
 * lead => pass illegal (rule_based/utils/utils.py:53-55); follow => pass legal (server/CFR.py:152-174) */
int ddz_ref_get_moves(const int8_t hand[15], const int8_t last[15], int8_t* out, int cap) {
    ensure();
    int lead = is_zero15(last), lc = 0, ll = 0, lv = 0, n = 0;
    if (!lead && ddz_ref_classify(last, &lc, &ll, &lv) < 0) return -2;
    for (int i = 0; i < UN; i++) {
        const act_t* a = &U[i];
        int ok = 1;
        for (int r = 0; r < 15; r++) if (a->cnt[r] > hand[r]) { ok = 0; break; }
        if (!ok) continue;
        if (lead) { if (a->cat == 0) continue; }
        else if (a->cat != 0 && !ddz_ref_bigger_than(a->cat, a->len, a->val, lc, ll, lv)) continue;
        if (n < cap && out) memcpy(out + 15 * n, a->cnt, 15);
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* constructive generator (same list, same order)                       */
/* ------------------------------------------------------------------ */
typedef struct { int8_t* out; int cap, n; } sink_t;
static inline void sink(sink_t* s, const int8_t c[15]) {
    if (s->out && s->n < s->cap) memcpy(s->out + 15 * s->n, c, 15);
    s->n++;
}
static void emit_main_kick(sink_t* s, int ms, int mL, int mmult, const int* kick, int k, int kmult) {
    int8_t c[15]; memset(c, 0, 15);
    for (int i = ms; i < ms + mL; i++) c[i] = (int8_t)mmult;
    for (int i = 0; i < k; i++) c[kick[i]] += (int8_t)kmult;
    sink(s, c);
}
typedef struct { sink_t* s; int ms, mL, mmult, kmult; } fk_ctx;
static void fk_cb(const int* pick, int k, void* v) {
    fk_ctx* f = (fk_ctx*)v; emit_main_kick(f->s, f->ms, f->mL, f->mmult, pick, k, f->kmult);
}
static int gen_fast(const int8_t hand[15], const int8_t last[15], sink_t* s) {
    ensure();
    int lead = is_zero15(last), lc = 0, ll = 0, lv = 0;
    if (!lead && ddz_ref_classify(last, &lc, &ll, &lv) < 0) return -2;
    if (!lead) { int8_t z[15]; memset(z, 0, 15); sink(s, z); }
    if (!lead && lc == 12) return s->n;
    /* per category: threshold count, line limits, kicker kind */
    static const int thr[15]  = {0, 1, 2, 3, 4, 3, 3, 1, 2, 3, 3, 3, 0, 4, 4};
    static const int lmin[15] = {0, 1, 1, 1, 1, 1, 1, 5, 3, 2, 2, 2, 0, 1, 1};
    static const int lmax[15] = {0, 1, 1, 1, 1, 1, 1, 12, 10, 6, 5, 4, 0, 1, 1};
    static const int kthr[15] = {0, 0, 0, 0, 0, 1, 2, 0, 0, 0, 1, 2, 0, 1, 2}; /* kicker needs count>=kthr */
    static const int knum[15] = {0, 0, 0, 0, 0, 1, 1, 0, 0, 0, -1, -1, 0, 2, 2}; /* -1: = len */
    for (int cat = 1; cat <= 14; cat++) {
        if (!lead && !(cat == lc || cat == 4 || cat == 12)) continue;
        if (cat == 12) {
            if (hand[13] && hand[14]) { int8_t c[15]; memset(c, 0, 15); c[13] = c[14] = 1; sink(s, c); }
            continue;
        }
        int same = (!lead && cat == lc);
        int smin = same ? lv + 1 : 0;
        int line = lmax[cat] > 1;
        int smax = line ? 11 : (thr[cat] == 1 ? 14 : 12);
        for (int st = smin; st <= smax; st++) {
            for (int L = lmin[cat]; L <= lmax[cat]; L++) {
                if (line && st + L > 12) break;
                int ok = 1;
                for (int i = st; i < st + L; i++) if (hand[i] < thr[cat]) { ok = 0; break; }
                if (!ok) break; /* longer runs from st fail too */
                if (same && L != ll) continue;
                int k = knum[cat] < 0 ? L : knum[cat];
                if (k == 0) { emit_main_kick(s, st, L, thr[cat], 0, 0, 0); continue; }
                int pool[15], n = 0;
                for (int r = 0; r < 15; r++)
                    if ((r < st || r >= st + L) && hand[r] >= kthr[cat]) pool[n++] = r;
                fk_ctx f = {s, st, L, thr[cat], kthr[cat]};
                combos(pool, n, k, fk_cb, &f);
            }
        }
    }
    return s->n;
}
int ddz_ref_get_moves_fast(const int8_t hand[15], const int8_t last[15], int8_t* out, int cap) {
    sink_t s = {out, cap, 0};
    return gen_fast(hand, last, &s);
}
int ddz_ref_count_moves_fast(const int8_t hand[15], const int8_t last[15]) {
    sink_t s = {0, 0, 0};
    return gen_fast(hand, last, &s);
}

/* ------------------------------------------------------------------ */
/* env                                                                  */
/* ------------------------------------------------------------------ */
void ddz_ref_env_clear(ddz_ref_env* e) { /* envi.py:30-36 + CEnv.reset */
    int32_t g = e->games;
    memset(e, 0, sizeof(*e));
    e->winner = -1; e->cur = 1; e->games = g;
}
static int rank_of(int id) { return id < 52 ? id / 4 : id - 39; } /* 52->13, 53->14 */

/* SURVEY App. C2 "reset": piles perm[0:17],[17:34],[34:51], bottom perm[51:54];
 * pile lord_pile (+bottom) -> role 1, next pile -> role 2, next -> role 0; lord moves first (game.py:173) */
int ddz_ref_env_deal(ddz_ref_env* e, const int8_t perm[54], int lord_pile) {
    if (lord_pile < 0 || lord_pile > 2) return -1;
    uint64_t seen = 0;
    for (int i = 0; i < 54; i++) { if (perm[i] < 0 || perm[i] > 53) return -1; seen |= 1ULL << perm[i]; }
    if (seen != (1ULL << 54) - 1) return -1;
    int32_t g = e->games;
    memset(e, 0, sizeof(*e));
    e->games = g + 1; e->winner = -1; e->cur = 1;
    static const int role_of_rel[3] = {1, 2, 0};
    for (int p = 0; p < 3; p++) {
        int role = role_of_rel[(p - lord_pile + 3) % 3];
        for (int i = 0; i < 17; i++) e->hand[role][rank_of(perm[17 * p + i])]++;
    }
    for (int i = 51; i < 54; i++) e->hand[1][rank_of(perm[i])]++;
    return 0;
}
static int left_of(const ddz_ref_env* e, int role) { int n = 0; for (int i = 0; i < 15; i++) n += e->hand[role][i]; return n; }

/* envi.py:103-110 (== server/core.py:60-62): previous player's hand-out, else the one before, else lead */
void ddz_ref_env_last(const ddz_ref_env* e, int8_t last[15]) {
    int p1 = (e->cur + 2) % 3, p2 = (e->cur + 1) % 3;
    if (!is_zero15(e->recent[p1])) memcpy(last, e->recent[p1], 15);
    else if (!is_zero15(e->recent[p2])) memcpy(last, e->recent[p2], 15);
    else memset(last, 0, 15);
}
int ddz_ref_env_legal(const ddz_ref_env* e, int8_t* out, int cap, int fast) {
    if (e->done) return 0;
    int8_t last[15]; ddz_ref_env_last(e, last);
    return fast ? ddz_ref_get_moves_fast(e->hand[e->cur], last, out, cap)
                : ddz_ref_get_moves(e->hand[e->cur], last, out, cap);
}

static void apply_move(ddz_ref_env* e, const int8_t move[15], int mcat, const int32_t R[3],
                       int* r, int* done, int* cat, float reward_out[3], int64_t* stats) {
    int s = e->cur, n = 0;
    for (int i = 0; i < 15; i++) {            /* envi.py:38-43 _update */
        e->hand[s][i] -= move[i]; e->hist[s][i] += move[i]; e->recent[s][i] = move[i]; n += move[i];
    }
    float rw[3] = {0, 0, 0};
    int rr = 0;
    if (n > 0 && left_of(e, s) == 0) {        /* rule_based/rule_play.py:14-28 */
        e->done = 1; e->winner = (int8_t)s; rr = (s == 1) ? -1 : 1;
        for (int q = 0; q < 3; q++) {         /* game.py:109-118: winners +R, losers -R; farmers share */
            int win = (s == 1) ? (q == 1) : (q != 1);
            rw[q] = win ? (float)R[q] : -(float)R[q];
        }
    }
    e->cur = (int8_t)((s + 1) % 3);           /* game.py:266 lord -> down -> up */
    if (r) *r = rr;
    if (done) *done = e->done;
    if (cat) *cat = mcat;
    if (reward_out) memcpy(reward_out, rw, sizeof rw);
    if (stats) {
        stats[4]++; if (n == 0) stats[9]++;
        if (e->done) {
            stats[0]++; stats[1] += (s == 1); stats[2] += (s == 2); stats[3] += (s == 0);
            stats[5] += (int64_t)rw[1]; stats[6] += (int64_t)rw[0] + (int64_t)rw[2];
        }
    }
}
static void noop_out(const ddz_ref_env* e, int* r, int* done, int* cat, float reward_out[3]) {
    if (r) *r = 0;
    if (done) *done = e->done;
    if (cat) *cat = -1;
    if (reward_out) reward_out[0] = reward_out[1] = reward_out[2] = 0.f;
}
int ddz_ref_env_step(ddz_ref_env* e, const int8_t move[15], const int32_t R[3],
                     int* r, int* done, int* cat, float reward_out[3]) {
    static const int32_t defR[3] = {50, 100, 50};
    if (!R) R = defR;
    if (e->done) { noop_out(e, r, done, cat, reward_out); return 0; }
    int8_t last[15]; ddz_ref_env_last(e, last);
    int mc = 0, ml = 0, mv = 0, legal = ddz_ref_classify(move, &mc, &ml, &mv) >= 0;
    if (legal) for (int i = 0; i < 15; i++) if (move[i] > e->hand[e->cur][i]) legal = 0;
    if (legal) {
        if (is_zero15(last)) legal = (mc != 0);
        else if (mc != 0) {
            int lc, ll, lv; ddz_ref_classify(last, &lc, &ll, &lv);
            legal = ddz_ref_bigger_than(mc, ml, mv, lc, ll, lv);
        }
    }
    if (!legal) { e->err = 1; noop_out(e, r, done, cat, reward_out); return -1; }
    apply_move(e, move, mc, R, r, done, cat, reward_out, 0);
    return 0;
}

/* SURVEY App. A get_state_prob_manual; inputs per server/core.py:26-33: known60 = thermometer one-hot of (played cards +
 * own hand), size1 / size2 = cards left of the next / next-next player; plane k = unknown60 * float(size_k)/float(size1+size2).
 * unknown60 is the one unpinned layout choice (oracle/SEMANTICS.md):
 *   form A (default)           thermometer(total - known): ones left-aligned like every other plane,
 *   form B (-DDDZ_PROB_FORM_B)  deck60 - known60 element by element (deck60 = 1111 x 13 ranks, 1000 per joker). */
static int total_of(int r) { return r < 13 ? 4 : 1; }
int ddz_ref_prob_form(void) {
#ifdef DDZ_PROB_FORM_B
    return 1;
#else
    return 0;
#endif
}
static void prob_planes60(const int known60[60], int size1, int size2, float out[120]) {
    int tot = size1 + size2;
    float p[2], unknown60[60];
    p[0] = tot > 0 ? (float)size1 / (float)tot : 0.f;
    p[1] = tot > 0 ? (float)size2 / (float)tot : 0.f;
    for (int r = 0; r < 15; r++) {
#ifdef DDZ_PROB_FORM_B
        for (int j = 0; j < 4; j++) {
            int deck = j < total_of(r) ? 1 : 0, d = deck - (known60[4 * r + j] != 0);
            unknown60[4 * r + j] = d > 0 ? 1.f : 0.f;
        }
#else
        int k = 0; for (int j = 0; j < 4; j++) k += known60[4 * r + j] != 0;
        int u = total_of(r) - k; if (u < 0) u = 0;
        for (int j = 0; j < 4; j++) unknown60[4 * r + j] = j < u ? 1.f : 0.f;
#endif
    }
    for (int k = 0; k < 2; k++) for (int i = 0; i < 60; i++) out[60 * k + i] = unknown60[i] * p[k];
}
void ddz_ref_state_prob_manual(const int32_t known60[60], int size1, int size2, float out[120]) {
    int k60[60];
    for (int i = 0; i < 60; i++) k60[i] = known60[i] != 0;
    prob_planes60(k60, size1, size2, out);
}
void ddz_ref_state_prob(const ddz_ref_env* e, float out[120]) {   /* envi.py:94: the same, from the env's own trackers */
    int known60[60], s = e->cur;
    for (int r = 0; r < 15; r++) {
        int known = e->hist[0][r] + e->hist[1][r] + e->hist[2][r] + e->hand[s][r];
        for (int j = 0; j < 4; j++) known60[4 * r + j] = j < known;
    }
    prob_planes60(known60, left_of(e, (s + 1) % 3), left_of(e, (s + 2) % 3), out);
}

static void thermo(const int* cnt, float* out) { /* envi.py:140-146: res[rank][:count] = 1 */
    for (int r = 0; r < 15; r++) for (int j = 0; j < 4; j++) out[4 * r + j] = (j < cnt[r]) ? 1.f : 0.f;
}
int ddz_ref_face_channels(int variant) {
    static const int C[4] = {4, 7, 9, 6};
    return (variant < 0 || variant > 3) ? -1 : C[variant];
}
/* envi.py:87-96 (Env), :165-178 (EnvComplicated), :182-198 (EnvCooperation), :202-217 (EnvCooperationSimplify) */
int ddz_ref_env_face(const ddz_ref_env* e, int variant, float* out) {
    int C = ddz_ref_face_channels(variant);
    if (C < 0) return -1;
    int s = e->cur, prev = (s + 2) % 3, next = (s + 1) % 3, prevprev = (s + 1) % 3;
    int planes[7][15], np = 0;
#define PL(expr) do { for (int r = 0; r < 15; r++) planes[np][r] = (expr); np++; } while (0)
    PL(e->hand[s][r]);
    PL(e->hist[0][r] + e->hist[1][r] + e->hist[2][r]);
    if (variant == 1 || variant == 2) { PL(e->hist[prev][r]); PL(e->hist[s][r]); PL(e->hist[next][r]); }
    if (variant == 2 || variant == 3) { PL(e->recent[prev][r]); PL(e->recent[prevprev][r]); }
#undef PL
    for (int i = 0; i < np; i++) thermo(planes[i], out + 60 * i);
    ddz_ref_state_prob(e, out + 60 * np);
    return C;
}
void ddz_ref_encode_actions(const int8_t* moves, int n, float* out) {
    for (int i = 0; i < n; i++) {
        int c[15]; for (int r = 0; r < 15; r++) c[r] = moves[15 * i + r];
        thermo(c, out + 60 * (size_t)i);
    }
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), counter (env_lo, env_hi, step, 0), key (seed_lo, seed_hi)        */
/* ------------------------------------------------------------------ */
uint32_t ddz_ref_philox(uint64_t seed, uint64_t env, uint32_t step) {
    uint32_t c0 = (uint32_t)env, c1 = (uint32_t)(env >> 32), c2 = step, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

/* ------------------------------------------------------------------ */
/* batched drivers                                                      */
/* ------------------------------------------------------------------ */
int64_t ddz_ref_batch_observe(const ddz_ref_env* envs, int B, int variant, int fast,
                              int32_t* offsets, uint64_t* actions_u64, float* actions_f32, int64_t cap,
                              float* face) {
    int C = ddz_ref_face_channels(variant);
    if (C < 0) return -1;
    int8_t buf[DDZ_REF_MAX_LEGAL * 15];
    int64_t total = 0;
    for (int b = 0; b < B; b++) {
        int n = ddz_ref_env_legal(&envs[b], buf, DDZ_REF_MAX_LEGAL, fast);
        if (n < 0 || n > DDZ_REF_MAX_LEGAL) return -2;
        if (offsets) offsets[b] = (int32_t)total;
        for (int i = 0; i < n; i++) {
            if (total + i >= cap) break;
            if (actions_u64) actions_u64[total + i] = ddz_ref_pack(buf + 15 * i);
            if (actions_f32) ddz_ref_encode_actions(buf + 15 * i, 1, actions_f32 + 60 * (total + i));
        }
        total += n;
        if (face) ddz_ref_env_face(&envs[b], variant, face + (size_t)b * C * 60);
    }
    if (offsets) offsets[B] = (int32_t)total;
    return total;
}

int ddz_ref_batch_step(ddz_ref_env* envs, int B, const int32_t* offsets, const uint64_t* actions_u64,
                       const int32_t* choice, int choice_mode, uint64_t seed, uint64_t env0, uint32_t step,
                       const int32_t rewards[3], int8_t* r, uint8_t* done, int8_t* cat, float* reward_out,
                       int64_t* stats) {
    static const int32_t defR[3] = {50, 100, 50};
    const int32_t* R = rewards ? rewards : defR;
    ensure();
    for (int b = 0; b < B; b++) {
        ddz_ref_env* e = &envs[b];
        int rr = 0, dd = e->done, cc = -1;
        float rw[3] = {0, 0, 0};
        if (!e->done) {
            int n = offsets[b + 1] - offsets[b];
            int64_t idx;
            if (choice_mode == 0) idx = choice[b];
            else if (choice_mode == 1) idx = n > 0 ? (int64_t)((uint32_t)choice[b] % (uint32_t)n) : -1;
            else idx = n > 0 ? (int64_t)(ddz_ref_philox(seed, env0 + (uint64_t)b, step) % (uint32_t)n) : -1;
            if (idx < 0 || idx >= n) { e->err = 1; if (stats) stats[7]++; }
            else {
                int8_t mv[15]; ddz_ref_unpack(actions_u64[offsets[b] + idx], mv);
                int mc, ml, mvv;
                if (ddz_ref_classify(mv, &mc, &ml, &mvv) < 0) { e->err = 1; if (stats) stats[7]++; }
                else apply_move(e, mv, mc, R, &rr, &dd, &cc, rw, stats);
            }
        }
        if (r) r[b] = (int8_t)rr;
        if (done) done[b] = (uint8_t)dd;
        if (cat) cat[b] = (int8_t)cc;
        if (reward_out) memcpy(reward_out + 3 * (size_t)b, rw, sizeof rw);
    }
    return 0;
}

int ddz_ref_batch_deal(ddz_ref_env* envs, int B, const int8_t* perm, const int8_t* lord_pile,
                       int only_done, int pool_games) {
    if (pool_games < 1) pool_games = 1;
    for (int b = 0; b < B; b++) {
        ddz_ref_env* e = &envs[b];
        if (only_done && !e->done) continue;
        size_t row = (size_t)(e->games % pool_games) * B + b;
        if (ddz_ref_env_deal(e, perm + 54 * row, lord_pile ? lord_pile[row] : 0) != 0) return -1;
    }
    return 0;
}

static void export_one(const ddz_ref_env* e, int B, int b, uint64_t* f, uint32_t* meta) {
    for (int q = 0; q < 3; q++) {
        f[(size_t)(0 + q) * B + b] = ddz_ref_pack(e->hand[q]);
        f[(size_t)(3 + q) * B + b] = ddz_ref_pack(e->hist[q]);
        f[(size_t)(6 + q) * B + b] = ddz_ref_pack(e->recent[q]);
    }
    uint32_t w = e->done ? (uint32_t)e->winner : 0u;
    meta[b] = (uint32_t)e->cur | ((uint32_t)e->done << 2) | (w << 3) | ((uint32_t)e->err << 5) |
              (((uint32_t)e->games & 0xFFFFFFu) << 8);
}
void ddz_ref_batch_export(const ddz_ref_env* envs, int B, uint64_t* f, uint32_t* meta) {
    for (int b = 0; b < B; b++) export_one(&envs[b], B, b, f, meta);
}

/* ------------------------------------------------------------------ */
/* CPU baseline rollout                                                 */
/* ------------------------------------------------------------------ */
typedef struct {
    int b0, b1, B, steps, warm, variant, pool_games;
    uint64_t seed, env0; const int8_t* perm; const int8_t* lord;
    int64_t stats[16]; uint64_t checksum; int64_t nsteps; double seconds;
    pthread_barrier_t* bar;
    uint64_t* out_fields; uint32_t* out_meta;   /* optional: the final state of every env (ddz_ref_rollout_export) */
} job_t;
static void export_one(const ddz_ref_env* e, int B, int b, uint64_t* f, uint32_t* meta);
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void* rollout_worker(void* arg) {
    job_t* j = (job_t*)arg;
    int C = ddz_ref_face_channels(j->variant);
    int n = j->b1 - j->b0;
    ddz_ref_env* envs = (ddz_ref_env*)calloc((size_t)n, sizeof(ddz_ref_env));
    int8_t* moves = (int8_t*)malloc(DDZ_REF_MAX_LEGAL * 15);
    float* af = (float*)malloc(sizeof(float) * 60 * DDZ_REF_MAX_LEGAL);
    float* face = (float*)malloc(sizeof(float) * 60 * C);
    static const int32_t R[3] = {50, 100, 50};
    uint64_t cs = 0;
    for (int i = 0; i < n; i++) {
        envs[i].winner = -1;
        size_t row = (size_t)(j->b0 + i);
        ddz_ref_env_deal(&envs[i], j->perm + 54 * row, j->lord ? j->lord[row] : 0);
    }
    double t0 = 0;
    for (int t = 0; t < j->warm + j->steps; t++) {
        if (t == j->warm) {   /* warm-up steps bring the envs to the steady-state mix; only the rest is timed */
            pthread_barrier_wait(j->bar);
            t0 = now_s();
            memset(j->stats, 0, sizeof j->stats); j->nsteps = 0;
        }
        for (int i = 0; i < n; i++) {
            ddz_ref_env* e = &envs[i];
            int b = j->b0 + i;
            int N = ddz_ref_env_legal(e, moves, DDZ_REF_MAX_LEGAL, 1);         /* a6 */
            ddz_ref_env_face(e, j->variant, face);                              /* a9 */
            ddz_ref_encode_actions(moves, N, af);                               /* a7 */
            j->stats[8] += N;
            uint32_t k = ddz_ref_philox(j->seed, j->env0 + (uint64_t)b, (uint32_t)t) % (uint32_t)N;
            int rr, dd, cc; float rw[3];
            int mc, ml, mv; ddz_ref_classify(moves + 15 * k, &mc, &ml, &mv);
            cs += ddz_ref_pack(moves + 15 * k) * 0x9E3779B97F4A7C15ULL + (uint64_t)N;
            { uint32_t u; memcpy(&u, &face[60 * (C - 2) + (t % 60)], 4); cs += u; memcpy(&u, &af[60 * k + 3], 4); cs += u; }
            apply_move(e, moves + 15 * k, mc, R, &rr, &dd, &cc, rw, j->stats); /* a10 */
            j->nsteps++;
            if (dd) {                                                           /* a2 re-deal */
                size_t row = (size_t)(e->games % j->pool_games) * j->B + b;
                ddz_ref_env_deal(e, j->perm + 54 * row, j->lord ? j->lord[row] : 0);
            }
        }
    }
    j->seconds = now_s() - t0;
    j->checksum = cs;
    if (j->out_fields && j->out_meta)
        for (int i = 0; i < n; i++) export_one(&envs[i], j->B, j->b0 + i, j->out_fields, j->out_meta);
    free(envs); free(moves); free(af); free(face);
    return 0;
}

static int64_t rollout_impl(int B, int warm_steps, int steps, int variant, uint64_t seed, uint64_t env0, const int8_t* perm_pool,
                            const int8_t* lord_pool, int pool_games, int nthreads, int64_t* stats,
                            uint64_t* checksum, double* seconds, uint64_t* out_fields, uint32_t* out_meta);
int64_t ddz_ref_rollout(int B, int warm_steps, int steps, int variant, uint64_t seed, const int8_t* perm_pool,
                        const int8_t* lord_pool, int pool_games, int nthreads, int64_t* stats,
                        uint64_t* checksum, double* seconds) {
    return rollout_impl(B, warm_steps, steps, variant, seed, 0, perm_pool, lord_pool, pool_games, nthreads, stats, checksum,
                        seconds, 0, 0);
}
/* the same rollout from the deal, returning the final state of EVERY env in the device's export layout
 * (ddz_ref_batch_export): the full-size parity check -- one diverging move anywhere changes some env's state */
int64_t ddz_ref_rollout_export(int B, int steps, int variant, uint64_t seed, uint64_t env0, const int8_t* perm_pool,
                               const int8_t* lord_pool, int pool_games, int nthreads, int64_t* stats,
                               uint64_t* fields9, uint32_t* meta) {
    if (!fields9 || !meta) return -1;
    return rollout_impl(B, 0, steps, variant, seed, env0, perm_pool, lord_pool, pool_games, nthreads, stats, 0, 0, fields9, meta);
}
static int64_t rollout_impl(int B, int warm_steps, int steps, int variant, uint64_t seed, uint64_t env0, const int8_t* perm_pool,
                            const int8_t* lord_pool, int pool_games, int nthreads, int64_t* stats,
                            uint64_t* checksum, double* seconds, uint64_t* out_fields, uint32_t* out_meta) {
    ensure();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
    if (pool_games < 1) pool_games = 1;
    if (ddz_ref_face_channels(variant) < 0) return -1;
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    pthread_barrier_t bar; pthread_barrier_init(&bar, 0, (unsigned)nthreads);
    if (warm_steps < 0) warm_steps = 0;
    for (int i = 0; i < nthreads; i++) {
        jobs[i].warm = warm_steps; jobs[i].bar = &bar;
        jobs[i].b0 = (int)((int64_t)B * i / nthreads); jobs[i].b1 = (int)((int64_t)B * (i + 1) / nthreads);
        jobs[i].B = B; jobs[i].steps = steps; jobs[i].variant = variant; jobs[i].pool_games = pool_games;
        jobs[i].seed = seed; jobs[i].env0 = env0; jobs[i].perm = perm_pool; jobs[i].lord = lord_pool;
        jobs[i].out_fields = out_fields; jobs[i].out_meta = out_meta;
        pthread_create(&th[i], 0, rollout_worker, &jobs[i]);
    }
    int64_t total = 0; uint64_t cs = 0; double sec = 0;
    if (stats) memset(stats, 0, 16 * sizeof(int64_t));
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], 0);
        total += jobs[i].nsteps; cs += jobs[i].checksum;
        if (jobs[i].seconds > sec) sec = jobs[i].seconds;
        if (stats) for (int k = 0; k < 16; k++) stats[k] += jobs[i].stats[k];
    }
    if (checksum) *checksum = cs;
    if (seconds) *seconds = sec;
    pthread_barrier_destroy(&bar);
    free(jobs); free(th);
    return total;
}

/* ------------------------------------------------------------------ */
/* CPU baseline of the legal-move microbench (BASELINE config 5)        */
/* ------------------------------------------------------------------ */
/* r.get_moves (envi.py:111) for n independent packed (hand, last) pairs, `reps` passes, split over nthreads; every list is
 * packed back to the product's uint64 format (that is the output a caller of the batched generator gets).  counts[n]
 * (optional) receives the list lengths, *checksum an order-independent digest of every emitted move. */
typedef struct { const uint64_t* hands; const uint64_t* lasts; int i0, i1, reps; int32_t* counts; int64_t moves;
                 uint64_t checksum; double seconds; pthread_barrier_t* bar; } gm_job;
static void* gm_worker(void* arg) {
    gm_job* j = (gm_job*)arg;
    int8_t* out = (int8_t*)malloc(DDZ_REF_MAX_LEGAL * 15);
    uint64_t cs = 0; int64_t moves = 0;
    pthread_barrier_wait(j->bar);
    double t0 = now_s();
    for (int rep = 0; rep < j->reps; rep++)
        for (int i = j->i0; i < j->i1; i++) {
            int8_t h[15], l[15];
            ddz_ref_unpack(j->hands[i], h); ddz_ref_unpack(j->lasts[i], l);
            int N = ddz_ref_get_moves_fast(h, l, out, DDZ_REF_MAX_LEGAL);
            if (N > DDZ_REF_MAX_LEGAL) N = DDZ_REF_MAX_LEGAL;
            for (int k = 0; k < N; k++) cs += ddz_ref_pack(out + 15 * k) * 0x9E3779B97F4A7C15ULL;
            if (j->counts && rep == 0) j->counts[i] = N;
            moves += N;
        }
    j->seconds = now_s() - t0; j->checksum = cs; j->moves = moves;
    free(out);
    return 0;
}
int64_t ddz_ref_get_moves_batch(const uint64_t* hands, const uint64_t* lasts, int n, int reps, int nthreads, int32_t* counts,
                                uint64_t* checksum, double* seconds) {
    ensure();
    if (n <= 0 || reps < 1) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n;
    gm_job* jobs = (gm_job*)calloc((size_t)nthreads, sizeof(gm_job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    pthread_barrier_t bar; pthread_barrier_init(&bar, 0, (unsigned)nthreads);
    for (int i = 0; i < nthreads; i++) {
        jobs[i].hands = hands; jobs[i].lasts = lasts; jobs[i].reps = reps; jobs[i].counts = counts; jobs[i].bar = &bar;
        jobs[i].i0 = (int)((int64_t)n * i / nthreads); jobs[i].i1 = (int)((int64_t)n * (i + 1) / nthreads);
        pthread_create(&th[i], 0, gm_worker, &jobs[i]);
    }
    int64_t total = 0; uint64_t cs = 0; double sec = 0;
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], 0);
        total += jobs[i].moves; cs += jobs[i].checksum;
        if (jobs[i].seconds > sec) sec = jobs[i].seconds;
    }
    if (checksum) *checksum = cs;
    if (seconds) *seconds = sec;
    pthread_barrier_destroy(&bar);
    free(jobs); free(th);
    return total;
}

/* digests of the r.get_moves lists of n packed pairs, for full-size comparisons: *unordered = sum over all moves of
 * pack(move) * phi, *ordered = sum of pack(move) * (2k + 1) * phi with k the move's index in ITS list (both mod 2^64);
 * counts[n] (optional) = list lengths.  Returns the number of moves. */
int64_t ddz_ref_get_moves_digest(const uint64_t* hands, const uint64_t* lasts, int n, int32_t* counts,
                                 uint64_t* unordered, uint64_t* ordered) {
    ensure();
    int8_t* out = (int8_t*)malloc(DDZ_REF_MAX_LEGAL * 15);
    uint64_t un = 0, ord = 0; int64_t moves = 0;
    for (int i = 0; i < n; i++) {
        int8_t h[15], l[15];
        ddz_ref_unpack(hands[i], h); ddz_ref_unpack(lasts[i], l);
        int N = ddz_ref_get_moves_fast(h, l, out, DDZ_REF_MAX_LEGAL);
        if (N < 0 || N > DDZ_REF_MAX_LEGAL) { free(out); return -1; }
        for (int k = 0; k < N; k++) {
            uint64_t p = ddz_ref_pack(out + 15 * k) * 0x9E3779B97F4A7C15ULL;
            un += p; ord += p * (uint64_t)(2 * k + 1);
        }
        if (counts) counts[i] = N;
        moves += N;
    }
    free(out);
    if (unordered) *unordered = un;
    if (ordered) *ordered = ord;
    return moves;
}

/* ------------------------------------------------------------------ */
/* exhaustive bound on the length of a legal-move list                  */
/* ------------------------------------------------------------------ */
/* Lead-move count of a hand in closed form (popcount x binomial), the same arithmetic the device uses; checked
 * against the enumerating generators in tests.  Used to PROVE DDZ_REF_MAX_LEGAL: the legal set is monotone in the
 * hand (containment), so the maximum over all hands is reached by a 20-card hand (the landlord's deal). */
static const int BINOM[16][6] = {
    {1, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0}, {1, 2, 1, 0, 0, 0}, {1, 3, 3, 1, 0, 0}, {1, 4, 6, 4, 1, 0},
    {1, 5, 10, 10, 5, 1}, {1, 6, 15, 20, 15, 6}, {1, 7, 21, 35, 35, 21}, {1, 8, 28, 56, 70, 56},
    {1, 9, 36, 84, 126, 126}, {1, 10, 45, 120, 210, 252}, {1, 11, 55, 165, 330, 462}, {1, 12, 66, 220, 495, 792},
    {1, 13, 78, 286, 715, 1287}, {1, 14, 91, 364, 1001, 2002}, {1, 15, 105, 455, 1365, 3003}};
static int binom_(int n, int k) { return (n < 0 || k > n || k > 5) ? 0 : BINOM[n][k]; }
static int lines_(unsigned R, int lmin, int lmax) {
    unsigned t = R; int n = 0;
    for (int L = 2; L <= lmax; L++) { t &= R >> (L - 1); if (L >= lmin) n += __builtin_popcount(t); }
    return n;
}
static int planes_(unsigned g3, int nk, int lmax) {
    unsigned R = g3 & 0xFFF, t = R; int n = 0;
    for (int L = 2; L <= lmax; L++) { t &= R >> (L - 1); n += __builtin_popcount(t) * binom_(nk - L, L); }
    return n;
}
int ddz_ref_count_lead_closed(const int8_t hand[15]) {
    unsigned g[5] = {0, 0, 0, 0, 0};
    for (int r = 0; r < 15; r++) for (int k = 1; k <= 4; k++) if (hand[r] >= k) g[k] |= 1u << r;
    if (!g[1]) return 0;
    int n1 = __builtin_popcount(g[1]), n2 = __builtin_popcount(g[2]), n3 = __builtin_popcount(g[3]), n4 = __builtin_popcount(g[4]);
    int n = n1 + n2 + n3 + n4 + n3 * (n1 - 1) + n3 * (n2 - 1);
    n += lines_(g[1] & 0xFFF, 5, 12) + lines_(g[2] & 0xFFF, 3, 10) + lines_(g[3] & 0xFFF, 2, 6);
    n += planes_(g[3], n1, 5) + planes_(g[3], n2, 4);
    n += ((g[1] & 0x6000) == 0x6000);
    n += n4 * binom_(n1 - 1, 2) + n4 * binom_(n2 - 1, 2);
    return n;
}
typedef struct { int first_lo, first_hi; int best; int8_t arg[15]; long long visited; } ex_job;
static void ex_rec(int8_t* h, int r, int left, ex_job* j) {
    if (r == 15) {
        if (left != 0) return;
        j->visited++;
        int c = ddz_ref_count_lead_closed(h);
        if (c > j->best) { j->best = c; memcpy(j->arg, h, 15); }
        return;
    }
    int cap = r < 13 ? 4 : 1;
    int rest = 0; for (int q = r + 1; q < 15; q++) rest += q < 13 ? 4 : 1;
    for (int c = 0; c <= cap && c <= left; c++) {
        if (left - c > rest) continue;
        h[r] = (int8_t)c; ex_rec(h, r + 1, left - c, j);
    }
    h[r] = 0;
}
static void* ex_worker(void* arg) {
    ex_job* j = (ex_job*)arg;
    int8_t h[15];
    for (int code = j->first_lo; code < j->first_hi; code++) {   /* code = counts of ranks 0 and 1 */
        memset(h, 0, 15);
        h[0] = (int8_t)(code / 5); h[1] = (int8_t)(code % 5);
        if (h[0] + h[1] <= 20) ex_rec(h, 2, 20 - h[0] - h[1], j);
    }
    return 0;
}
/* max number of lead moves over ALL 20-card hands of one deck; argmax hand in best_hand; *visited = hands examined */
int ddz_ref_max_lead_moves_exhaustive(int nthreads, int8_t best_hand[15], long long* visited) {
    ensure();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 25) nthreads = 25;
    ex_job jobs[25]; pthread_t th[25];
    for (int i = 0; i < nthreads; i++) {
        memset(&jobs[i], 0, sizeof jobs[i]);
        jobs[i].first_lo = 25 * i / nthreads; jobs[i].first_hi = 25 * (i + 1) / nthreads;
        pthread_create(&th[i], 0, ex_worker, &jobs[i]);
    }
    int best = 0; long long vis = 0;
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], 0);
        vis += jobs[i].visited;
        if (jobs[i].best > best) { best = jobs[i].best; if (best_hand) memcpy(best_hand, jobs[i].arg, 15); }
    }
    if (visited) *visited = vis;
    return best;
}

/* ------------------------------------------------------------------ */
/* the search bot's move list (server/mcts/evaluator.py, get_moves.py)  */
/* ------------------------------------------------------------------ */
/* the cards of a universe entry in the order card.py:34-159 writes them (rank indices 0..14): mains first, ascending,
 * then the kickers ascending -- and for pair kickers the kicker LIST twice (`list(extra) * 2`, card.py:130,154) */
static int action_cards(const act_t* a, int* out) {
    int n = 0, main_mult = 0;
    switch (a->cat) {
        case 0: return 0;
        case 1: case 2: case 3: case 4: main_mult = a->cat; break;
        case 5: case 6: case 9: case 10: case 11: main_mult = 3; break;
        case 7: main_mult = 1; break;
        case 8: main_mult = 2; break;
        case 12: out[0] = 13; out[1] = 14; return 2;
        default: main_mult = 4; break;                       /* 13, 14 */
    }
    int L = (a->cat >= 7 && a->cat <= 11) ? a->len : 1;
    for (int r = a->val; r < a->val + L; r++) for (int i = 0; i < main_mult; i++) out[n++] = r;
    int kmult = (a->cat == 5 || a->cat == 10 || a->cat == 13) ? 1 : (a->cat == 6 || a->cat == 11 || a->cat == 14) ? 2 : 0;
    if (kmult) {
        int kick[8], nk = 0;
        for (int r = 0; r < 15; r++) if ((r < a->val || r >= a->val + L) && a->cnt[r]) kick[nk++] = r;
        if (a->cat == 6) { out[n++] = kick[0]; out[n++] = kick[0]; }          /* card.py:82  [extra] * 2 */
        else for (int rep = 0; rep < kmult; rep++) for (int i = 0; i < nk; i++) out[n++] = kick[i];
    }
    return n;
}

/* evaluator.py:17-57, branch for branch; char2val = rank index + 3 (evaluator.py:4-9) */
double ddz_ref_cards_value(const int8_t counts[15]) {
    ensure();
    for (int i = 0; i < 15; i++) if (counts[i] < 0 || counts[i] > 4) return -1e30;
    int u = lookup(ddz_ref_pack(counts));
    if (u < 0 || U[u].extra) return -1e30;                     /* not a key of cards_value: KeyError in the reference */
    const act_t* act = &U[u];
    int a[24], n = action_cards(act, a), c = act->cat;
#define VAL(i) (a[i] + 3)
    double v = 0;
    if (c == 0) v = 0;
    else if (c <= 3) {
        v = VAL(0) - 10;
        if (c == 2 && v > 0) v *= 1.5;
        if (c == 3 && v > 0) v *= 2;
    } else if (c == 4) v = 9;
    else if (c <= 6) {
        v = VAL(0) - 10;
        if (v > 0) v *= 1.5;
    } else if (c <= 9) {
        v = (VAL(n - 1) - 10) / 2.0; if (v < 0) v = 0;
    } else if (c == 10) {
        int main_len = n / 4 * 3;
        v = (VAL(n - 1) - 10) / 2.0; if (v < 0) v = 0;
        for (int i = main_len; i < n; i++) if (VAL(i) > 10) v += VAL(i) - 10;
    } else if (c == 11) {
        int main_len = n / 5 * 3;
        v = (VAL(n - 1) - 10) / 2.0; if (v < 0) v = 0;
        for (int i = main_len; i < main_len + main_len / 3; i++) if (VAL(i) > 10) v += 1.5 * (VAL(i) - 10);
    } else if (c == 12) v = 12;
    else v = VAL(0) - 10;
#undef VAL
    return v;
}

/* get_moves.py:36-69: the r.get_moves list, pruned when it has more than 10 entries -- the rocket-kicker moves go
 * (:22-34,:56-57), the rest is ranked by  cards_value - 0.1 * (cards left in hand after the move)  with Python's stable
 * sort (:60), and the list becomes  lowest, highest, 2nd lowest, 2nd highest, ...  for int(length/3 + 1) rounds (:61-63),
 * `length` counting the dropped moves too.  An index past the kept moves (IndexError in the reference; cannot happen
 * for hands of one deck) is clamped.  Returns the number of entries (only cap are written). */
int ddz_ref_mcts_moves(const int8_t hand[15], const int8_t last[15], int8_t* out, int cap) {
    int8_t all[DDZ_REF_MAX_LEGAL * 15];
    int length = ddz_ref_get_moves_fast(hand, last, all, DDZ_REF_MAX_LEGAL);
    if (length < 0 || length > DDZ_REF_MAX_LEGAL) return -1;
    if (length <= 10) {
        for (int i = 0; i < length && i < cap; i++) memcpy(out + 15 * i, all + 15 * i, 15);
        return length;
    }
    int handnum = 0;
    for (int r = 0; r < 15; r++) handnum += hand[r];
    int kept[DDZ_REF_MAX_LEGAL], order[DDZ_REF_MAX_LEGAL], m = 0;
    double values[DDZ_REF_MAX_LEGAL];
    for (int i = 0; i < length; i++) {
        const int8_t* mv = all + 15 * i;
        int u = lookup(ddz_ref_pack(mv));
        if (u < 0) return -2;
        if (U[u].extra) continue;                              /* sidaihuojian / sandaihuojian */
        int cards = 0;
        for (int r = 0; r < 15; r++) cards += mv[r];
        volatile double left = 0.1 * (handnum - cards);        /* two roundings, as Python does: no fused multiply-add */
        values[m] = ddz_ref_cards_value(mv) - left;
        kept[m] = i; order[m] = m; m++;
    }
    for (int i = 1; i < m; i++) {                              /* stable insertion sort by value */
        int o = order[i], j = i;
        while (j > 0 && values[order[j - 1]] > values[o]) { order[j] = order[j - 1]; j--; }
        order[j] = o;
    }
    if (m == 0) return 0;                                      /* nothing left to rank (cannot happen: a bomb or the rocket stays) */
    int rounds = length / 3 + 1, n = 0;
    for (int k = 0; k < rounds; k++) {
        int lo = k < m ? k : m - 1, hi = m - 1 - k >= 0 ? m - 1 - k : 0;
        if (n < cap) memcpy(out + 15 * n, all + 15 * kept[order[lo]], 15);
        n++;
        if (n < cap) memcpy(out + 15 * n, all + 15 * kept[order[hi]], 15);
        n++;
    }
    return n;
}
