/*
 * oracle/sdx_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the SuperDendrix
 * significance-and-search path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's
 * library; the product (superdendrix_b200/) never does.
 *
 * Every function cites the reference lines (relative to /root/reference/) it
 * restates.  Where the reference draws from CPython's Mersenne Twister
 * (random.sample / random.shuffle) the restatement draws from the canonical
 * counter-based stream defined in DESIGN.md §3 (Philox4x32-10), because the
 * MT stream is not a parity target; the *distribution* and the
 * integer outcome of every trade under a replayed schedule are.
 *
 * Pinning: this oracle is pinned by tests/test_oracle_golden.py against
 * fixtures produced by importing the reference itself
 * (tests/golden/make_golden.py): replayed Curveball schedules, SuperW values,
 * brute-force scan optima, scipy rank-sums.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SDXO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants) */
/* ------------------------------------------------------------------ */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

SDXO_API void sdxo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

/* A sequential view of one counter-based stream:
 *   counter = (block, t, id_lo, (id_hi & 0x3fffffff) | (kind << 30)), key = seed.
 * Draw i of the stream is word (i & 3) of block (i >> 2). */
enum { STREAM_PAIR = 0, STREAM_DEAL = 1, STREAM_PERM = 2 };

typedef struct {
    uint32_t key[2];
    uint32_t t, id_lo, id_hi_kind;
    uint32_t block;
    uint32_t buf[4];
    int pos;
} stream_t;

static void stream_init(stream_t *s, uint64_t seed, int kind, uint64_t id, uint32_t t) {
    s->key[0] = (uint32_t)seed;
    s->key[1] = (uint32_t)(seed >> 32);
    s->t = t;
    s->id_lo = (uint32_t)id;
    s->id_hi_kind = ((uint32_t)(id >> 32) & 0x3fffffffu) | ((uint32_t)kind << 30);
    s->block = 0;
    s->pos = 4;
}

static uint32_t stream_next(stream_t *s) {
    if (s->pos == 4) {
        uint32_t ctr[4] = {s->block, s->t, s->id_lo, s->id_hi_kind};
        sdxo_philox4x32_10(ctr, s->key, s->buf);
        s->block++;
        s->pos = 0;
    }
    return s->buf[s->pos++];
}

/* Exactly uniform integer in [0, m), m >= 1 (Lemire 2019, with rejection).
 * Plays the role of CPython's _randbelow under random.sample / shuffle
 * (call sites src/curveball.py:22,32, src/superdendrix.py:373). */
static uint32_t stream_below(stream_t *s, uint32_t m) {
    uint64_t prod = (uint64_t)stream_next(s) * m;
    uint32_t lo = (uint32_t)prod;
    if (lo < m) {
        uint32_t thr = (uint32_t)(-m) % m;
        while (lo < thr) {
            prod = (uint64_t)stream_next(s) * m;
            lo = (uint32_t)prod;
        }
    }
    return (uint32_t)(prod >> 32);
}

SDXO_API void sdxo_stream_draws(uint64_t seed, int kind, uint64_t id, uint32_t t, int n, uint32_t *out) {
    stream_t s;
    stream_init(&s, seed, kind, id, t);
    for (int i = 0; i < n; ++i) out[i] = stream_next(&s);
}

/* test hook for the rejection branch: returns the value and how many draws it used */
SDXO_API uint32_t sdxo_below_from_draws(const uint32_t *draws, int n_draws, uint32_t m, int *used) {
    int i = 0;
    uint64_t prod = (uint64_t)draws[i++] * m;
    uint32_t lo = (uint32_t)prod;
    if (lo < m) {
        uint32_t thr = (uint32_t)(-m) % m;
        while (lo < thr && i < n_draws) {
            prod = (uint64_t)draws[i++] * m;
            lo = (uint32_t)prod;
        }
    }
    *used = i;
    return (uint32_t)(prod >> 32);
}

/* ------------------------------------------------------------------ */
/* bit helpers; bit j of word i of a row = element 32*i + j            */
/* ------------------------------------------------------------------ */
static inline int popc32(uint32_t x) { return __builtin_popcount(x); }

/* position (0..31) of the d-th (0-based) set bit of x */
static inline int nth_set_bit(uint32_t x, int d) {
    while (d-- > 0) x &= x - 1;
    return __builtin_ctz(x);
}

/* ------------------------------------------------------------------ */
/* Curveball                                                           */
/* ------------------------------------------------------------------ */

/* Trade pair of trade t of null `null_id`: an ordered pair of two distinct
 * rows, uniform over the R(R-1) possibilities — random.sample(range(R), 2)
 * at src/curveball.py:22-24. */
SDXO_API void sdxo_trade_pair(uint64_t seed, uint64_t null_id, uint32_t t, uint32_t R, uint32_t *a, uint32_t *b) {
    stream_t s;
    stream_init(&s, seed, STREAM_PAIR, null_id, t);
    uint32_t x = stream_below(&s, R);
    uint32_t y = stream_below(&s, R - 1);
    if (y >= x) y++;
    *a = x; *b = y;
}

/* Draw `attempt` of pick i of trade t of null `null_id` (DEAL stream):
 * word (i & 3) of the Philox block with counter
 *   ((i >> 2) | (attempt << 16), t, id_lo, (id_hi & 0x3fffffff) | (DEAL << 30)).
 * Every pick owns its own sub-stream, so the picks of one trade can be drawn
 * in parallel and a (rare) Lemire rejection of one pick does not shift the
 * draws of the others. */
static uint32_t deal_draw(uint64_t seed, uint64_t null_id, uint32_t t, uint32_t i, uint32_t attempt) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(i >> 2) | (attempt << 16), t, (uint32_t)null_id,
                       ((uint32_t)(null_id >> 32) & 0x3fffffffu) | ((uint32_t)STREAM_DEAL << 30)};
    uint32_t out[4];
    sdxo_philox4x32_10(ctr, key, out);
    return out[i & 3];
}

/* how often the rejection zone of Lemire's method was entered / a draw was rejected (test instrumentation) */
static uint64_t g_deal_zone = 0, g_deal_redraw = 0;
SDXO_API void sdxo_deal_stats(uint64_t *zone, uint64_t *redraw, int reset) {
    if (zone) *zone = g_deal_zone;
    if (redraw) *redraw = g_deal_redraw;
    if (reset) g_deal_zone = g_deal_redraw = 0;
}

/* exactly uniform integer in [0, m) for pick i (Lemire with rejection) */
static uint32_t deal_below(uint64_t seed, uint64_t null_id, uint32_t t, uint32_t i, uint32_t m) {
    uint32_t attempt = 0;
    uint64_t prod = (uint64_t)deal_draw(seed, null_id, t, i, attempt) * m;
    uint32_t lo = (uint32_t)prod;
    if (lo < m) {
        uint32_t thr = (uint32_t)(-m) % m;
        g_deal_zone++;
        while (lo < thr) {
            g_deal_redraw++;
            prod = (uint64_t)deal_draw(seed, null_id, t, i, ++attempt) * m;
            lo = (uint32_t)prod;
        }
    }
    return (uint32_t)(prod >> 32);
}

SDXO_API uint32_t sdxo_deal_draw(uint64_t seed, uint64_t null_id, uint32_t t, uint32_t i, uint32_t attempt) {
    return deal_draw(seed, null_id, t, i, attempt);
}

/* One canonical trade on packed rows ra, rb (wpr words each).
 * Restates src/curveball.py:25-35: keep the common part, skip when one row
 * contains the other (:29), re-deal the symmetric difference so that both row
 * sizes are preserved (:30-35).  The uniformly random |a\b|-subset that the
 * reference obtains by shuffle(tot)[:L] (:32-34) is obtained here as a
 * uniformly random k-subset, k = min(|a\b|, |b\a|), for the smaller side:
 * the n symmetric-difference elements are ranked 0..n-1 in ascending bit
 * order and the ranks are chosen with Floyd's algorithm
 *     for i in 0..k-1:  j = n-k+i;  r = below_i(j+1);  chosen += (r in chosen ? j : r)
 * which yields every k-subset with equal probability.
 * `chosen` is scratch of at least wpr+1 words.  Returns 1 if the trade was applied. */
static int trade_canonical(uint32_t *ra, uint32_t *rb, int wpr, uint64_t seed, uint64_t null_id, uint32_t t,
                           uint32_t *chosen) {
    int la = 0, lb = 0;
    for (int i = 0; i < wpr; ++i) {
        la += popc32(ra[i] & ~rb[i]);
        lb += popc32(rb[i] & ~ra[i]);
    }
    if (la == 0 || lb == 0) return 0;           /* l_ab in [l_a, l_b]  (:29) */
    int n = la + lb;
    int small_is_a = la <= lb;
    int k = small_is_a ? la : lb;
    memset(chosen, 0, sizeof(uint32_t) * ((size_t)wpr + 1));
    for (int i = 0; i < k; ++i) {
        uint32_t j = (uint32_t)(n - k + i);
        uint32_t r = deal_below(seed, null_id, t, (uint32_t)i, j + 1);
        if ((chosen[r >> 5] >> (r & 31)) & 1u) r = j;
        chosen[r >> 5] |= 1u << (r & 31);
    }
    uint32_t rank = 0;
    for (int i = 0; i < wpr; ++i) {
        uint32_t ab = ra[i] & rb[i], sym = ra[i] ^ rb[i], pick = 0;
        for (uint32_t x = sym; x; x &= x - 1, ++rank)
            if ((chosen[rank >> 5] >> (rank & 31)) & 1u) pick |= x & (0u - x);
        if (small_is_a) { ra[i] = ab | pick; rb[i] = ab | (sym ^ pick); }
        else            { rb[i] = ab | pick; ra[i] = ab | (sym ^ pick); }
    }
    return 1;
}

/* ------------------------------------------------------------------ */
/* Burn-in of the pool states                          */
/* ------------------------------------------------------------------ */
/* Pair of a BURN-IN trade when the burn orientation still has dominant rows: with probability 1/4 both rows come from
 * the K densest rows (top[], by size descending, index ascending), otherwise the pair is uniform as in sdxo_trade_pair.
 * Row sizes are invariants of the chain, so this is a fixed pair distribution and the stationary distribution stays
 * uniform.  K < 2: the plain uniform pair of sdxo_trade_pair (same draws). */
static void trade_pair_burn(uint64_t seed, uint64_t null_id, uint32_t t, uint32_t R, const uint16_t *top, uint32_t K,
                            uint32_t *a, uint32_t *b) {
    if (K < 2) { sdxo_trade_pair(seed, null_id, t, R, a, b); return; }
    stream_t s;
    stream_init(&s, seed, STREAM_PAIR, null_id, t);
    uint32_t u = stream_next(&s);
    if (u < 0x40000000u) {
        uint32_t i = stream_below(&s, K), j = stream_below(&s, K - 1);
        if (j >= i) j++;
        *a = top[i]; *b = top[j];
    } else {
        uint32_t x = stream_below(&s, R), y = stream_below(&s, R - 1);
        if (y >= x) y++;
        *a = x; *b = y;
    }
}

/* null_id: the stream the elements are dealt from (the round's own id); pair_id: the stream the row pairs are drawn from */
static int64_t curveball_burn_round(uint32_t *rows, int R, int wpr, uint64_t seed, uint64_t null_id, uint64_t pair_id, int64_t n_trades,
                                    const uint16_t *top, uint32_t K) {
    uint32_t *chosen = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)wpr + 1));
    int64_t applied = 0;
    for (int64_t t = 0; t < n_trades; ++t) {
        uint32_t a, b;
        trade_pair_burn(seed, pair_id, (uint32_t)t, (uint32_t)R, top, K, &a, &b);
        applied += trade_canonical(rows + (size_t)a * wpr, rows + (size_t)b * wpr, wpr, seed, null_id, (uint32_t)t, chosen);
    }
    free(chosen);
    return applied;
}

/* the K = max(2, ceil(R / 10)) densest rows, by size descending then index ascending */
static uint32_t dense_rows(const int *size, int R, uint16_t *top) {
    uint32_t K = (uint32_t)((R + 9) / 10);
    if (K < 2) K = 2;
    if (K > (uint32_t)R) K = (uint32_t)R;
    char *used = (char *)calloc((size_t)R, 1);
    for (uint32_t q = 0; q < K; ++q) {
        int best = -1;
        for (int r = 0; r < R; ++r)
            if (!used[r] && (best < 0 || size[r] > size[best])) best = r;
        used[best] = 1;
        top[q] = (uint16_t)best;
    }
    free(used);
    return K;
}

/* out (n_cols rows x wpr_out words) = bit transpose of in (n_rows rows x wpr_in words) */
static void transpose_bits(const uint32_t *in, int n_rows, int wpr_in, int n_cols, uint32_t *out, int wpr_out) {
    memset(out, 0, sizeof(uint32_t) * (size_t)n_cols * wpr_out);
    for (int r = 0; r < n_rows; ++r)
        for (int c = 0; c < n_cols; ++c)
            if ((in[(size_t)r * wpr_in + (c >> 5)] >> (c & 31)) & 1u) out[(size_t)c * wpr_out + (r >> 5)] |= 1u << (r & 31);
}

/* How the pool states of one matrix are burnt in.  Curveball may trade the rows of either orientation of
 * the matrix: both chains keep both margins and have the uniform distribution as their stationary distribution.  What
 * mixes slowly is a row far larger than its partners (it exchanges at most |partner| elements per trade), so the
 * burn-in runs in the orientation whose largest row is the smaller multiple of the mean row:
 *     transposed (rows over the LONGER dimension)  iff  L <= 12000  and  max_col * L < max_row * R
 * (R x L = the trade orientation of src/curveball.py:10-11).  A burn-in round is 5 * R_b trades (at most 60000), R_b the
 * rows of that orientation; its pairs are uniform unless even there the largest row exceeds 8 x the mean row
 * (max_b * R_b > 8 * ones), in which case they favour the densest tenth of the rows (trade_pair_burn). */
typedef struct {
    int alt;                /* 1: rows over the longer dimension */
    int Rb, Lb, wprb;       /* rows, bits per row, words per row of the burn orientation */
    int64_t Tb;             /* trades per burn-in round */
    uint32_t *obs;          /* the observed matrix in the burn orientation */
    uint16_t *top;
    uint32_t K;             /* 0: uniform pairs */
} burn_plan;

static void burn_plan_init(burn_plan *bp, const uint32_t *observed, int R, int wpr, int L) {
    int *rs = (int *)calloc((size_t)R, sizeof(int)), *cs = (int *)calloc((size_t)L, sizeof(int));
    int64_t ones = 0;
    int max_r = 0, max_c = 0;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < L; ++c)
            if ((observed[(size_t)r * wpr + (c >> 5)] >> (c & 31)) & 1u) { rs[r]++; cs[c]++; ones++; }
    for (int r = 0; r < R; ++r) if (rs[r] > max_r) max_r = rs[r];
    for (int c = 0; c < L; ++c) if (cs[c] > max_c) max_c = cs[c];
    bp->alt = L <= 12000 && (int64_t)max_c * L < (int64_t)max_r * R;
    if (bp->alt) {
        bp->Rb = L; bp->Lb = R; bp->wprb = (R + 31) / 32;
        bp->obs = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)bp->Rb * bp->wprb);
        transpose_bits(observed, R, wpr, L, bp->obs, bp->wprb);
    } else {
        bp->Rb = R; bp->Lb = L; bp->wprb = wpr;
        bp->obs = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)R * wpr);
        memcpy(bp->obs, observed, sizeof(uint32_t) * (size_t)R * wpr);
    }
    bp->Tb = 5 * (int64_t)bp->Rb;
    if (bp->Tb > 60000) bp->Tb = 60000;
    const int *size = bp->alt ? cs : rs;
    const int max_b = bp->alt ? max_c : max_r;
    bp->top = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)bp->Rb);
    bp->K = ((int64_t)max_b * bp->Rb > 8 * ones) ? dense_rows(size, bp->Rb, bp->top) : 0;
    free(rs); free(cs);
}

static void burn_plan_free(burn_plan *bp) { free(bp->obs); free(bp->top); }

/* start state j of the pool, in the trade orientation (R x wpr): `burn` rounds from the observed matrix on the reserved
 * stream ids (B + j) * burn + r, r < burn, with B = 2^61 / burn rounded UP to a multiple of 32.  Like the nulls (DESIGN.md
 * §3c) the states of an aligned block of 32 share their schedule of row pairs: round r of state j draws its pairs from the
 * stream of round r of state j - j % 32 and deals the elements from its own stream. */
static void burn_state(const burn_plan *bp, int R, int wpr, uint64_t seed, int64_t burn, int j, uint32_t *out) {
    size_t wb = (size_t)bp->Rb * bp->wprb;
    uint32_t *st = (uint32_t *)malloc(sizeof(uint32_t) * wb);
    memcpy(st, bp->obs, sizeof(uint32_t) * wb);
    uint64_t B = (((1ull << 61) / (uint64_t)burn + 31) / 32) * 32;
    uint64_t base = (B + (uint64_t)j) * (uint64_t)burn, lead = (B + (uint64_t)(j - j % 32)) * (uint64_t)burn;
    for (int64_t r = 0; r < burn; ++r)
        curveball_burn_round(st, bp->Rb, bp->wprb, seed, base + (uint64_t)r, lead + (uint64_t)r, bp->Tb, bp->top, bp->K);
    if (bp->alt) transpose_bits(st, bp->Rb, bp->wprb, R, out, wpr);
    else memcpy(out, st, sizeof(uint32_t) * wb);
    free(st);
}

/* what burn_plan_init decided, for the tests: info = {alt, Rb, Tb, K} */
SDXO_API void sdxo_burn_plan(const uint32_t *observed, int R, int wpr, int L, int64_t *info) {
    burn_plan bp;
    burn_plan_init(&bp, observed, R, wpr, L);
    info[0] = bp.alt; info[1] = bp.Rb; info[2] = bp.Tb; info[3] = bp.K;
    burn_plan_free(&bp);
}

/* curve_ball(input_matrix, r_hp, num_iterations) — src/curveball.py:17-40 —
 * on the packed form of r_hp (R rows over the longer dimension), in place,
 * for null index `null_id`.  Returns the number of applied trades. */
/* pair_id: the stream the row pairs are drawn from — the null's own index, or, when blocks of chains
 * share their schedule of row pairs (pair_group > 1), the index of the same round of the block's first chain.  The
 * elements are always dealt from the null's own DEAL stream. */
static int64_t curveball_round(uint32_t *rows, int R, int wpr, uint64_t seed, uint64_t null_id, uint64_t pair_id, int64_t n_trades) {
    uint32_t *chosen = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)wpr + 1));
    int64_t applied = 0;
    for (int64_t t = 0; t < n_trades; ++t) {
        uint32_t a, b;
        sdxo_trade_pair(seed, pair_id, (uint32_t)t, (uint32_t)R, &a, &b);
        applied += trade_canonical(rows + (size_t)a * wpr, rows + (size_t)b * wpr, wpr, seed, null_id, (uint32_t)t, chosen);
    }
    free(chosen);
    return applied;
}

SDXO_API int64_t sdxo_curveball(uint32_t *rows, int R, int wpr, uint64_t seed, uint64_t null_id, int64_t n_trades) {
    return curveball_round(rows, R, wpr, seed, null_id, null_id, n_trades);
}

/* stream id of the row pairs of null `id`: round j of chain c uses the pairs of round j of chain c - c % pair_group */
static uint64_t pair_stream_id(uint64_t id, int64_t chain_len, int pair_group) {
    if (pair_group <= 1) return id;
    uint64_t c = id / (uint64_t)chain_len, j = id % (uint64_t)chain_len;
    return (c - c % (uint64_t)pair_group) * (uint64_t)chain_len + j;
}

/* The null-driver loop of src/generate_null_matrices.py:71-74: nulls
 * null0 .. null0+n-1; null i starts a new chain when i % chain_len == 0 and
 * continues from null i-1 otherwise (chain_len = n reproduces the reference's
 * single Markov chain over the mutable `presence`, chain_len = 1 the
 * independent restarts of src/curveball_main.py:136).
 * Burn-in: with burn > 0 a chain does not start from the
 * observed matrix but from start state (c % pool_size) of its chain index c,
 * where state j = the observed matrix after `burn` burn-in rounds (burn_state;
 * L = bits per row is needed there: the burn-in may run on the transpose).
 * out = n packed matrices (may be NULL); applied[n] optional. */
SDXO_API void sdxo_curveball_nulls_group(const uint32_t *observed, int R, int wpr, int L, uint64_t seed,
                                         uint64_t null0, int64_t n, int64_t n_trades, int64_t chain_len,
                                         int64_t burn, int pool_size, int pair_group, uint32_t *out, int64_t *applied) {
    size_t msz = (size_t)R * wpr;
    uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * msz);
    uint32_t **pool = NULL;
    burn_plan bp;
    if (burn > 0) {
        pool = (uint32_t **)calloc((size_t)pool_size, sizeof(uint32_t *));
        burn_plan_init(&bp, observed, R, wpr, L);
    }
    for (int64_t i = 0; i < n; ++i) {
        uint64_t id = null0 + (uint64_t)i;
        if (i == 0 || id % (uint64_t)chain_len == 0) {
            uint64_t chain = id / (uint64_t)chain_len;
            if (burn > 0) {
                int j = (int)(chain % (uint64_t)pool_size);
                if (!pool[j]) {
                    pool[j] = (uint32_t *)malloc(sizeof(uint32_t) * msz);
                    burn_state(&bp, R, wpr, seed, burn, j, pool[j]);
                }
                memcpy(cur, pool[j], sizeof(uint32_t) * msz);
            } else {
                memcpy(cur, observed, sizeof(uint32_t) * msz);
            }
            /* a shard that starts inside a chain replays the chain prefix */
            for (uint64_t q = chain * (uint64_t)chain_len; q < id; ++q)
                curveball_round(cur, R, wpr, seed, q, pair_stream_id(q, chain_len, pair_group), n_trades);
        }
        int64_t ap = curveball_round(cur, R, wpr, seed, id, pair_stream_id(id, chain_len, pair_group), n_trades);
        if (applied) applied[i] = ap;
        if (out) memcpy(out + (size_t)i * msz, cur, sizeof(uint32_t) * msz);
    }
    if (pool) {
        for (int j = 0; j < pool_size; ++j) free(pool[j]);
        free(pool);
        burn_plan_free(&bp);
    }
    free(cur);
}

/* start state j of the burn-in pool, as the chains of sdxo_curveball_nulls_group use it */
SDXO_API void sdxo_pool_state(const uint32_t *observed, int R, int wpr, int L, uint64_t seed, int64_t burn, int j,
                              uint32_t *out) {
    burn_plan bp;
    burn_plan_init(&bp, observed, R, wpr, L);
    burn_state(&bp, R, wpr, seed, burn, j, out);
    burn_plan_free(&bp);
}

SDXO_API void sdxo_curveball_nulls(const uint32_t *observed, int R, int wpr, uint64_t seed,
                                   uint64_t null0, int64_t n, int64_t n_trades, int64_t chain_len,
                                   uint32_t *out, int64_t *applied) {
    sdxo_curveball_nulls_group(observed, R, wpr, wpr * 32, seed, null0, n, n_trades, chain_len, 0, 0, 1, out, applied);
}

/* Replay of a recorded reference schedule: trade t uses
 * rows a[t], b[t]; to_a[t] (wpr words) holds the symmetric-difference elements
 * the reference dealt to row a (tot[:L], src/curveball.py:33-34).  Skipped
 * trades (:29) are detected from the rows themselves.  Returns #applied. */
SDXO_API int64_t sdxo_curveball_replay(uint32_t *rows, int R, int wpr, const int32_t *a, const int32_t *b,
                                       const uint32_t *to_a, int64_t n_trades) {
    (void)R;
    int64_t applied = 0;
    for (int64_t t = 0; t < n_trades; ++t) {
        uint32_t *ra = rows + (size_t)a[t] * wpr, *rb = rows + (size_t)b[t] * wpr;
        const uint32_t *m = to_a + (size_t)t * wpr;
        int la = 0, lb = 0;
        for (int i = 0; i < wpr; ++i) {
            la += popc32(ra[i] & ~rb[i]);
            lb += popc32(rb[i] & ~ra[i]);
        }
        if (la == 0 || lb == 0) continue;
        for (int i = 0; i < wpr; ++i) {
            uint32_t ab = ra[i] & rb[i], sym = ra[i] ^ rb[i];
            ra[i] = ab | (sym & m[i]);
            rb[i] = ab | (sym & ~m[i]);
        }
        applied++;
    }
    return applied;
}

/* ------------------------------------------------------------------ */
/* Set objective                                                       */
/* ------------------------------------------------------------------ */

/* SuperW — src/superdendrix.py:387-391 — for the feature rows feats[0..k) of a
 * packed F x S matrix, weights w[S] (0 for samples outside the profile) and an
 * optional sample mask (NULL = all samples).  Also the integer outputs of
 * src/superdendrix.py:229-236,345-350: coverage |U|, overlap sum|g| - |U|,
 * act_cov |U ∩ {w>0}|. */
SDXO_API void sdxo_score_set(const uint32_t *rows, int wpr, int S, const int32_t *feats, int k,
                             const double *w, const uint32_t *mask,
                             double *W, int32_t *cov, int32_t *overlap, int32_t *act_cov) {
    double pos = 0.0, neg = 0.0;
    int c = 0, tot = 0, ac = 0;
    for (int s = 0; s < S; ++s) {
        int wi = s >> 5;
        uint32_t bit = 1u << (s & 31);
        if (mask && !(mask[wi] & bit)) continue;
        int mult = 0;
        for (int j = 0; j < k; ++j)
            if (feats[j] >= 0 && (rows[(size_t)feats[j] * wpr + wi] & bit)) mult++;
        if (!mult) continue;
        c++; tot += mult;
        if (w[s] > 0) { pos += w[s]; ac++; }
        neg += mult * fabs(w[s]);
    }
    *W = 2.0 * pos - neg;
    *cov = c; *overlap = tot - c; *act_cov = ac;
}

/* Per-feature outputs — src/superdendrix.py:281-286 (sum of w over carriers,
 * first maximum wins), :219-227 (W of the singleton), :229-232 (|g|). */
SDXO_API void sdxo_feature_scores(const uint32_t *rows, int F, int wpr, int S, const double *w,
                                  const uint32_t *mask, double *sum_w, double *W1, int32_t *count, int32_t *best) {
    double zbest = -INFINITY; int ebest = -1;
    for (int g = 0; g < F; ++g) {
        double sw = 0, pos = 0, neg = 0; int c = 0;
        for (int s = 0; s < S; ++s) {
            uint32_t bit = 1u << (s & 31);
            if (mask && !(mask[s >> 5] & bit)) continue;
            if (!(rows[(size_t)g * wpr + (s >> 5)] & bit)) continue;
            c++; sw += w[s];
            if (w[s] > 0) pos += w[s];
            neg += fabs(w[s]);
        }
        sum_w[g] = sw; W1[g] = 2 * pos - neg; count[g] = c;
        if (sw > zbest) { zbest = sw; ebest = g; }
    }
    if (best) *best = ebest;
}

/* ------------------------------------------------------------------ */
/* Exhaustive scan: max over |M| <= k of W(M) — the integral optimum of the
 * ILP of src/superdendrix.py:294-313,393-401.  Enumeration
 * order = size-major, then lexicographic; ties keep the earlier subset.
 * topn best subsets are returned sorted by (W desc, enumeration rank asc). */
/* ------------------------------------------------------------------ */
typedef struct { double W; int32_t f[3]; int32_t cov; int64_t rank; } sdxo_hit;

static void hit_insert(sdxo_hit *top, int *n, int cap, const sdxo_hit *h) {
    int i = *n;
    if (i == cap) {
        if (!(h->W > top[cap - 1].W)) return;
        i = cap - 1;
    } else (*n)++;
    while (i > 0 && h->W > top[i - 1].W) { top[i] = top[i - 1]; --i; }
    top[i] = *h;
}

/* Subsets whose smallest feature index lies in [a_lo, a_hi) (the unit the scan is sharded by); a_lo = 0, a_hi = F is
 * the whole enumeration.  Returns the number of subsets scored.  Ties keep enumeration order (size-major, then
 * lexicographic) within the range. */
SDXO_API int64_t sdxo_scan_range(const uint32_t *rows, int F, int wpr, int S, const double *w, const uint32_t *mask,
                                 int k, int topn, int a_lo, int a_hi, double *outW, int32_t *outF, int32_t *outCov) {
    sdxo_hit *top = (sdxo_hit *)calloc((size_t)topn, sizeof(sdxo_hit));
    int n = 0; int64_t rank = 0;
    int32_t f[3]; int32_t ov, ac; sdxo_hit h;
    for (int size = 1; size <= k && size <= 3; ++size) {
        for (int a = a_lo; a < a_hi && a < F; ++a) {
            if (size == 1) {
                f[0] = a; f[1] = f[2] = -1;
                sdxo_score_set(rows, wpr, S, f, 1, w, mask, &h.W, &h.cov, &ov, &ac);
                memcpy(h.f, f, sizeof f); h.rank = rank++; hit_insert(top, &n, topn, &h);
                continue;
            }
            for (int b = a + 1; b < F; ++b) {
                if (size == 2) {
                    f[0] = a; f[1] = b; f[2] = -1;
                    sdxo_score_set(rows, wpr, S, f, 2, w, mask, &h.W, &h.cov, &ov, &ac);
                    memcpy(h.f, f, sizeof f); h.rank = rank++; hit_insert(top, &n, topn, &h);
                    continue;
                }
                for (int c = b + 1; c < F; ++c) {
                    f[0] = a; f[1] = b; f[2] = c;
                    sdxo_score_set(rows, wpr, S, f, 3, w, mask, &h.W, &h.cov, &ov, &ac);
                    memcpy(h.f, f, sizeof f); h.rank = rank++; hit_insert(top, &n, topn, &h);
                }
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        outW[i] = top[i].W; outCov[i] = top[i].cov;
        outF[3 * i] = top[i].f[0]; outF[3 * i + 1] = top[i].f[1]; outF[3 * i + 2] = top[i].f[2];
    }
    for (int i = n; i < topn; ++i) { outW[i] = -INFINITY; outCov[i] = 0; outF[3*i] = outF[3*i+1] = outF[3*i+2] = -1; }
    free(top);
    return rank;
}

SDXO_API int64_t sdxo_scan(const uint32_t *rows, int F, int wpr, int S, const double *w, const uint32_t *mask,
                           int k, int topn, double *outW, int32_t *outF, int32_t *outCov) {
    return sdxo_scan_range(rows, F, wpr, S, w, mask, k, topn, 0, F, outW, outF, outCov);
}

/* ------------------------------------------------------------------ */
/* Conditional permutation test — src/superdendrix.py:361-384.
 * For slot i of the module: P times replace the carrier set of feature i by a
 * uniform random |g_i|-subset of the profile's samples (:373) and count
 * W >= real_W (:374).  With U = union of the other features,
 *   W = W(M \ g_i) + sum_{s in R} v_s,  v_s = 2 w_s [w_s>0 and s notin U] - |w_s|
 * (SURVEY.md §8 a7).  The subset is drawn with Floyd's algorithm from stream
 * (PERM, id = perm_id, t = permutation index); sums accumulate in draw order.
 * sample_idx[n_samples] lists the profile's samples (ascending). */
/* ------------------------------------------------------------------ */
SDXO_API void sdxo_condperm(const uint32_t *rows, int wpr, int S, const int32_t *feats, int k,
                            const double *w, const int32_t *sample_idx, int n_samples,
                            uint64_t seed, uint64_t perm_id0, int64_t P, int64_t *counts, double *real_W_out) {
    int nw = (S + 31) / 32;
    uint32_t *U = (uint32_t *)calloc((size_t)nw * 2, sizeof(uint32_t));
    uint32_t *chosen = U + nw;
    uint32_t *maskw = (uint32_t *)calloc((size_t)nw, sizeof(uint32_t));
    for (int i = 0; i < n_samples; ++i) maskw[sample_idx[i] >> 5] |= 1u << (sample_idx[i] & 31);
    double *v = (double *)malloc(sizeof(double) * (size_t)n_samples);
    for (int i = 0; i < k; ++i) {
        /* W of the module without slot i, and U = union of the others */
        int32_t others[8]; int ko = 0;
        for (int j = 0; j < k; ++j) if (j != i) others[ko++] = feats[j];
        double base; int32_t c, ov, ac;
        sdxo_score_set(rows, wpr, S, others, ko, w, maskw, &base, &c, &ov, &ac);
        memset(U, 0, sizeof(uint32_t) * nw);
        for (int j = 0; j < ko; ++j)
            for (int q = 0; q < nw; ++q) U[q] |= rows[(size_t)others[j] * wpr + q];
        int n_i = 0; double real_sum = 0.0;
        for (int q = 0; q < n_samples; ++q) {
            int s = sample_idx[q];
            int inU = (U[s >> 5] >> (s & 31)) & 1;
            v[q] = ((w[s] > 0 && !inU) ? 2.0 * w[s] : 0.0) - fabs(w[s]);
            if ((rows[(size_t)feats[i] * wpr + (s >> 5)] >> (s & 31)) & 1) { n_i++; real_sum += v[q]; }
        }
        double real_W = base + real_sum;
        if (real_W_out) real_W_out[i] = real_W;
        int64_t cnt = 0;
        for (int64_t p = 0; p < P; ++p) {
            stream_t st;
            stream_init(&st, seed, STREAM_PERM, perm_id0 + (uint64_t)i, (uint32_t)p);
            memset(chosen, 0, sizeof(uint32_t) * nw);
            double sum = 0.0;
            for (int j = n_samples - n_i; j < n_samples; ++j) {
                uint32_t t = stream_below(&st, (uint32_t)j + 1);
                if ((chosen[t >> 5] >> (t & 31)) & 1) t = (uint32_t)j;
                chosen[t >> 5] |= 1u << (t & 31);
                sum += v[t];
            }
            cnt += (base + sum >= real_W);
        }
        counts[i] = cnt;
    }
    free(U); free(maskw); free(v);
}

/* ------------------------------------------------------------------ */
/* Wilcoxon rank-sum as scipy.stats.ranksums computes it (call sites
 * utils/compute_p_value.py:306-318, utils/compute_ranksum_pval.py:116-128):
 * mid-ranks of the pooled sample, no tie or continuity correction.
 * x = profile over carriers of the union, y = over the other samples.
 * Also returns 2*sum(rank(x)) as an exact integer. */
/* ------------------------------------------------------------------ */
typedef struct { double v; int idx; } vi_t;
static int cmp_vi(const void *a, const void *b) {
    double x = ((const vi_t *)a)->v, y = ((const vi_t *)b)->v;
    return (x > y) - (x < y);
}

SDXO_API void sdxo_ranksum(const double *profile, const int32_t *sample_idx, int n_samples,
                           const uint32_t *union_bits, double *z, double *p, int64_t *twice_rank_sum, int32_t *n1_out) {
    vi_t *a = (vi_t *)malloc(sizeof(vi_t) * (size_t)n_samples);
    for (int i = 0; i < n_samples; ++i) { a[i].v = profile[sample_idx[i]]; a[i].idx = sample_idx[i]; }
    qsort(a, (size_t)n_samples, sizeof(vi_t), cmp_vi);
    int64_t s2 = 0; int n1 = 0;
    for (int i = 0; i < n_samples;) {
        int j = i;
        while (j < n_samples && a[j].v == a[i].v) ++j;
        int64_t twice_mid = (int64_t)(i + 1) + (int64_t)j;    /* 2 * mean of ranks i+1..j */
        for (int q = i; q < j; ++q) {
            int s = a[q].idx;
            if ((union_bits[s >> 5] >> (s & 31)) & 1) { s2 += twice_mid; n1++; }
        }
        i = j;
    }
    int n2 = n_samples - n1;
    double expected = n1 * (double)(n1 + n2 + 1) / 2.0;
    double zz = (0.5 * (double)s2 - expected) / sqrt((double)n1 * n2 * (n1 + n2 + 1) / 12.0);
    *z = zz;
    *p = erfc(fabs(zz) / sqrt(2.0));
    *twice_rank_sum = s2; *n1_out = n1;
    free(a);
}
