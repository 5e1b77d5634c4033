/*
 * TEST INFRASTRUCTURE ONLY -- oracle/kmer_oracle.c
 *
 * Plain-C CPU restatement of the reference k-mer index (Clemapfel/kmer_index) build + search path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker. The product (libkmer_b200.so) never links, calls or falls
 * back to anything in here.
 *
 * Parity pinning: the reference ships NO golden vectors or known-answer files for this path
 * (its only test, test_main.cpp, is a randomized differential against seqan3::fm_index, which is
 * absent here). This restatement is therefore pinned against the reference ITSELF, compiled from
 * /root/reference into oracle/_ref/libkmer_ref.so (oracle/ref_driver.cpp, oracle/Makefile), by
 * tests/test_oracle_vs_reference.py, and against the fixtures under tests/golden/ that were
 * generated from that compiled reference (tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it restates. Deliberate representation changes (all
 * result-preserving, all exercised by the pinning tests):
 *   - `_data` (robin_hood::unordered_map<hash, vector<pos>>, kmer_index.hpp:52) is held as one array of
 *     positions stably sorted by hash plus the sorted hash array; a bucket is the run of equal hashes.
 *     Within a bucket positions ascend, exactly as push_back in text order produces (kmer_index.hpp:165).
 *   - the sigma^(k-size) map probes of get_position_for_all_kmer_with_prefix (kmer_index.hpp:138-144)
 *     become the contiguous run of hashes in [lower_bound, upper_bound) of that sorted array, and the
 *     per-bucket std::binary_search for `pos + k` (kmer_index.hpp:240-247) becomes "the k-mer starting
 *     at pos+k has its hash inside [lower_bound, upper_bound)", recomputed from a retained text copy
 *     -- the same predicate, since position x sits in bucket hash(T[x..x+k)) and nowhere else.
 *   - undefined behaviour `*it` with it == end() (kmer_index.hpp:317,546) is evaluated as "not equal".
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KO_OK 0
#define KO_THROW_INVALID_ARGUMENT 1 /* std::invalid_argument: kmer_index.hpp:121 and :508 */
#define KO_UNDEFINED 2              /* reference behaviour undefined (m == 0, m == 10000) */

#define KO_UB_FLAG 0x80             /* or-ed into status: the reference dereferenced an end() iterator
                                       (kmer_index.hpp:317,546) while answering this query; what it
                                       returns then depends on stale heap bytes. The restatement evaluates
                                       the comparison as "not equal"; pinning tests skip flagged queries. */

#define KO_QUERY_SIZE_RANGE 10000 /* kmer_index.hpp:401 */

static __thread int ko_ub_seen;

/* ---------------------------------------------------------------- fast_pow.hpp:46-93 */
uint64_t ko_fast_pow(uint64_t base, uint8_t exp)
{
    /* highest_bit_set[exp] (fast_pow.hpp:10-44): number of significant bits of exp for exp < 63,
       255 ("overflow") for exp >= 63 */
    if (exp >= 63)
        return base == 1 ? 1 : 0; /* fast_pow.hpp:52-60 */
    uint64_t result = 1;
    /* the fall-through switch (fast_pow.hpp:62-91) is square-and-multiply over the bits of exp,
       arithmetic mod 2^64 */
    while (exp) {
        if (exp & 1)
            result *= base;
        exp >>= 1;
        base *= base;
    }
    return result;
}

/* ---------------------------------------------------------------- kmer_index.hpp:56-73 */
uint64_t ko_hash(const uint8_t *ranks, uint32_t k, uint32_t sigma)
{
    uint64_t h = 0;
    for (uint32_t i = 0; i < k; ++i)
        h += (uint64_t)ranks[i] * ko_fast_pow(sigma, (uint8_t)(k - i - 1));
    return h;
}

/* ---------------------------------------------------------------- index storage */
typedef struct {
    uint32_t k;
    uint64_t n_kmers;  /* n - k + 1 */
    uint64_t *hashes;  /* sorted ascending, one per k-mer start */
    uint32_t *pos;     /* positions, stably sorted by hash  (== concatenated buckets of _data) */
    uint8_t *last_kmer; /* kmer_index.hpp:87,174 */
    uint64_t *needed;   /* restricted index only (ko_create_restricted): bit h set <=> bucket h is held completely */
} ko_element;

/* restricted index: number of at()/range lookups that asked for a bucket the index does not hold (must stay 0) */
static volatile uint64_t ko_restricted_misses_;

typedef struct {
    uint32_t sigma;
    uint64_t n;
    uint8_t *text;
    uint32_t n_ks;
    uint32_t ks[64];      /* template order */
    uint32_t all_ks[64];  /* _all_ks after the descending sort, kmer_index.hpp:410 */
    ko_element *elems;    /* template order */
    /* choose_search_scheme tables, kmer_index.hpp:404-405 */
    uint64_t *sum_off;    /* [KO_QUERY_SIZE_RANGE + 1] */
    uint8_t *sum_ks;      /* flattened _optimal_nk_sum */
    uint8_t *use_multi;   /* _use_multi_search_scheme */
} ko_index;

/* stable LSD radix sort of (hash, pos) by hash */
static void sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t n, uint64_t max_key)
{
    uint64_t *k2 = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    uint32_t *v2 = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint64_t *cnt = (uint64_t *)malloc(65536 * sizeof(uint64_t));
    for (int shift = 0; shift < 64 && (max_key >> shift) != 0; shift += 16) {
        memset(cnt, 0, 65536 * sizeof(uint64_t));
        for (uint64_t i = 0; i < n; ++i)
            cnt[(keys[i] >> shift) & 0xFFFF]++;
        uint64_t s = 0;
        for (int b = 0; b < 65536; ++b) {
            uint64_t c = cnt[b];
            cnt[b] = s;
            s += c;
        }
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t d = cnt[(keys[i] >> shift) & 0xFFFF]++;
            k2[d] = keys[i];
            v2[d] = vals[i];
        }
        memcpy(keys, k2, n * sizeof(uint64_t));
        memcpy(vals, v2, n * sizeof(uint32_t));
    }
    free(k2);
    free(v2);
    free(cnt);
}

/* kmer_index.hpp:154-179  create(text) */
static int element_create(ko_element *e, uint32_t k, const uint8_t *text, uint64_t n, uint32_t sigma)
{
    e->k = k;
    e->n_kmers = n - k + 1;
    e->hashes = (uint64_t *)malloc(e->n_kmers * sizeof(uint64_t));
    e->pos = (uint32_t *)malloc(e->n_kmers * sizeof(uint32_t));
    e->last_kmer = (uint8_t *)malloc(k);
    if (!e->hashes || !e->pos || !e->last_kmer)
        return -1;
    /* :157  text | views::kmer_hash(ungapped{k}) -- rolling evaluation of the same polynomial */
    uint64_t top = ko_fast_pow(sigma, (uint8_t)(k - 1));
    uint64_t h = ko_hash(text, k, sigma);
    uint64_t max_key = h;
    e->hashes[0] = h;
    e->pos[0] = 0;
    for (uint64_t p = 1; p < e->n_kmers; ++p) {
        h = (h - (uint64_t)text[p - 1] * top) * sigma + text[p + k - 1];
        e->hashes[p] = h;
        e->pos[p] = (uint32_t)p; /* :165  _data[h].push_back(i) */
        if (h > max_key)
            max_key = h;
    }
    sort_pairs(e->hashes, e->pos, e->n_kmers, max_key);
    /* :174  _last_kmer = text[n-k, n);  :177-178  _last_kmer_refs[j] = { n - k + j } (implicit) */
    memcpy(e->last_kmer, text + n - k, k);
    return 0;
}

typedef struct {
    const uint32_t *p;
    uint64_t len;
} ko_bucket;

/* kmer_index.hpp:76-84  at(hash): pointer to the bucket or null */
static int element_at(const ko_element *e, uint64_t hash, ko_bucket *out)
{
    if (e->needed && !((e->needed[hash >> 6] >> (hash & 63)) & 1))
        __atomic_add_fetch(&ko_restricted_misses_, 1, __ATOMIC_RELAXED);
    uint64_t lo = 0, hi = e->n_kmers;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (e->hashes[mid] < hash)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo == e->n_kmers || e->hashes[lo] != hash)
        return 0;
    uint64_t a = lo, b = e->n_kmers;
    uint64_t l2 = lo;
    while (l2 < b) {
        uint64_t mid = l2 + (b - l2) / 2;
        if (e->hashes[mid] <= hash)
            l2 = mid + 1;
        else
            b = mid;
    }
    out->p = e->pos + a;
    out->len = l2 - a;
    return 1;
}

static uint64_t lower_bound_hash(const ko_element *e, uint64_t hash)
{
    uint64_t lo = 0, hi = e->n_kmers;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (e->hashes[mid] < hash)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* std::lower_bound over a bucket */
static uint64_t bucket_lower_bound(const ko_bucket *b, uint64_t value)
{
    uint64_t lo = 0, hi = b->len;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if ((uint64_t)b->p[mid] < value)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* std::binary_search over a bucket */
static int bucket_binary_search(const ko_bucket *b, uint64_t value)
{
    uint64_t i = bucket_lower_bound(b, value);
    return i < b->len && (uint64_t)b->p[i] == value;
}

/* ---------------------------------------------------------------- result (kmer_index_result.hpp) */
typedef struct {
    ko_bucket *buckets; /* _positions: pointers into the index (kmer_index_result.hpp:22-23) */
    uint64_t n_buckets, cap_buckets;
    uint8_t *mask; /* _bitmask, one byte per bit for simplicity (compressed_bitset.hpp) */
    uint64_t n_mask;
    int bypass; /* BYPASS_BITMASK, kmer_index_result.hpp:11 */
    uint32_t tail[64]; /* storage for the singleton buckets _last_kmer_refs[i] referenced by a sub-k result */
    uint32_t n_tail;
} ko_result;

static void result_init(ko_result *r)
{
    memset(r, 0, sizeof(*r));
}

static void result_free(ko_result *r)
{
    free(r->buckets);
    free(r->mask);
    result_init(r);
}

static void result_push(ko_result *r, const uint32_t *p, uint64_t len)
{
    if (r->n_buckets == r->cap_buckets) {
        r->cap_buckets = r->cap_buckets ? 2 * r->cap_buckets : 8;
        r->buckets = (ko_bucket *)realloc(r->buckets, r->cap_buckets * sizeof(ko_bucket));
    }
    r->buckets[r->n_buckets].p = p;
    r->buckets[r->n_buckets].len = len;
    r->n_buckets++;
}

/* kmer_index_result.hpp:211-217: one bucket + bitmask filled with zero_or_one (or bypassed) */
static void result_single(ko_result *r, const ko_bucket *b, int fill, int bypass)
{
    result_push(r, b->p, b->len);
    r->bypass = bypass;
    if (!bypass) {
        r->n_mask = b->len;
        r->mask = (uint8_t *)malloc(b->len ? b->len : 1);
        memset(r->mask, fill ? 1 : 0, b->len ? b->len : 1);
    }
}

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/* kmer_index_result.hpp:244-260  to_vector(): concatenate, filter by bitmask, std::sort */
static uint64_t result_to_vector(const ko_result *r, uint32_t **out)
{
    uint64_t total = 0;
    for (uint64_t b = 0; b < r->n_buckets; ++b)
        total += r->buckets[b].len;
    uint32_t *v = (uint32_t *)malloc((total ? total : 1) * sizeof(uint32_t));
    uint64_t o = 0, i = 0;
    for (uint64_t b = 0; b < r->n_buckets; ++b)
        for (uint64_t j = 0; j < r->buckets[b].len; ++j, ++i)
            if (r->bypass || (i < r->n_mask && r->mask[i]))
                v[o++] = r->buckets[b].p[j];
    int sorted = 1;
    for (uint64_t j = 1; j < o && sorted; ++j)
        sorted = v[j - 1] <= v[j];
    if (!sorted)
        qsort(v, o, sizeof(uint32_t), cmp_u32);
    *out = v;
    return o;
}

/* ---------------------------------------------------------------- element search */
typedef struct {
    uint64_t lower, upper; /* hash range [lower, upper)  kmer_index.hpp:131-133 */
    uint8_t tail_match[64]; /* check_last_kmer: tail_match[i] <=> _last_kmer[i, i+size) == prefix */
    uint32_t size;
} ko_prefix;

/* kmer_index.hpp:115-148 get_position_for_all_kmer_with_prefix (+ check_last_kmer :90-112).
   Returns KO_THROW_INVALID_ARGUMENT for the :119-122 throw. */
static int element_prefix(const ko_index *ix, const ko_element *e, const uint8_t *prefix, uint32_t size, ko_prefix *out)
{
    const uint32_t k = e->k;
    if ((double)ko_fast_pow(ix->sigma, (uint8_t)(k - size)) > 1e7) /* :119 */
        return KO_THROW_INVALID_ARGUMENT;
    uint64_t prefix_hash = 0;
    for (uint32_t i = 0; i < size; ++i) /* :126-129 */
        prefix_hash += (uint64_t)prefix[i] * ko_fast_pow(ix->sigma, (uint8_t)(k - i - 1));
    out->lower = prefix_hash;
    out->upper = prefix_hash + ko_fast_pow(ix->sigma, (uint8_t)(k - size));
    out->size = size;
    memset(out->tail_match, 0, sizeof(out->tail_match));
    for (uint32_t i = 1; i < k - size + 1; ++i) { /* :94-111 */
        int equal = 1;
        for (uint32_t j = i; j < i + size; ++j)
            if (e->last_kmer[j] != prefix[j - i]) {
                equal = 0;
                break;
            }
        out->tail_match[i] = (uint8_t)equal;
    }
    return KO_OK;
}

/* "x is contained in one of the buckets returned by get_position_for_all_kmer_with_prefix"
   (the any-of over std::binary_search at kmer_index.hpp:240-247) */
static int prefix_contains(const ko_index *ix, const ko_element *e, const ko_prefix *pf, uint64_t x)
{
    const uint32_t k = e->k;
    if (x + k <= ix->n) { /* x is a k-mer start: it lives in bucket hash(T[x, x+k)) */
        uint64_t h = ko_hash(ix->text + x, k, ix->sigma);
        return h >= pf->lower && h < pf->upper;
    }
    /* _last_kmer_refs[i] = { n - k + i }, kmer_index.hpp:177-178 */
    if (x < ix->n) {
        uint64_t i = x - (ix->n - k);
        if (i >= 1 && i < (uint64_t)k - pf->size + 1)
            return pf->tail_match[i];
    }
    return 0;
}

/* kmer_index.hpp:193-346  kmer_index_element::search */
static int element_search(const ko_index *ix, const ko_element *e, const uint8_t *q, uint64_t m, ko_result *res)
{
    const uint32_t k = e->k;
    if (m == k) { /* :198-205 */
        ko_bucket b;
        if (element_at(e, ko_hash(q, k, ix->sigma), &b))
            result_single(res, &b, 1, 1);
        return KO_OK;
    }
    if (m > k) { /* :207-339 */
        uint64_t rest_n = m % k;
        uint64_t n_parts = (m - rest_n) / k;
        ko_bucket *nk = (ko_bucket *)malloc(n_parts * sizeof(ko_bucket));
        for (uint64_t j = 0; j < n_parts; ++j) { /* :216-227 (the last_hash cache only skips a lookup) */
            if (!element_at(e, ko_hash(q + j * k, k, ix->sigma), &nk[j])) {
                free(nk);
                return KO_OK; /* :224 empty result */
            }
        }
        const ko_bucket *back = &nk[n_parts - 1];
        uint8_t *usable = (uint8_t *)malloc(back->len ? back->len : 1); /* :230 */
        memset(usable, 1, back->len ? back->len : 1);
        if (rest_n > 0) { /* :232-256 */
            ko_prefix pf;
            int st = element_prefix(ix, e, q + m - rest_n, (uint32_t)rest_n, &pf);
            if (st != KO_OK) {
                free(nk);
                free(usable);
                return st;
            }
            for (uint64_t i = 0; i < back->len; ++i)
                usable[i] = (uint8_t)prefix_contains(ix, e, &pf, (uint64_t)back->p[i] + k);
        }
        if (n_parts == 1) { /* :259-267 */
            result_single(res, &nk[0], 0, 0);
            for (uint64_t i = 0; i < nk[0].len; ++i)
                if (usable[i])
                    res->mask[i] = 1;
        } else if (rest_n == 0) { /* :270-298 */
            result_single(res, &nk[0], 1, 0);
            for (uint64_t s = 0; s < nk[0].len; ++s) {
                uint64_t previous_pos = nk[0].p[s];
                int should_use = 1;
                for (uint64_t j = 1; j < n_parts; ++j) {
                    if (!bucket_binary_search(&nk[j], previous_pos + k)) {
                        res->mask[s] = 0;
                        should_use = 0;
                        break;
                    } else
                        previous_pos += k;
                }
                if (should_use)
                    res->mask[s] = 1;
            }
        } else { /* :301-338  -- note `current = nk_positions.back()` at :314 for EVERY later part */
            result_single(res, &nk[0], 1, 0);
            for (uint64_t s = 0; s < nk[0].len; ++s) {
                uint64_t previous_pos = nk[0].p[s];
                int interrupted = 0;
                for (uint64_t j = 1; j < n_parts; ++j) {
                    const ko_bucket *current = back;
                    uint64_t it = bucket_lower_bound(current, previous_pos += k);
                    if (it == current->len)
                        ko_ub_seen = 1;
                    if (it == current->len || (uint64_t)current->p[it] != previous_pos) { /* :317 */
                        interrupted = 1;
                        break;
                    }
                    if (j == n_parts - 1) { /* :323-329 */
                        if (!usable[it])
                            interrupted = 1;
                        break;
                    }
                }
                if (interrupted)
                    res->mask[s] = 0;
            }
        }
        free(nk);
        free(usable);
        return KO_OK;
    }
    /* m < k : :342-345  result_t(get_position_for_all_kmer_with_prefix(query.begin(), query.size())) */
    ko_prefix pf;
    int st = element_prefix(ix, e, q, (uint32_t)m, &pf);
    if (st != KO_OK)
        return st;
    if (e->needed && !(((e->needed[pf.lower >> 6] >> (pf.lower & 63)) & 1)))
        __atomic_add_fetch(&ko_restricted_misses_, 1, __ATOMIC_RELAXED);
    uint64_t lo = lower_bound_hash(e, pf.lower), hi = lower_bound_hash(e, pf.upper);
    if (hi > lo)
        result_push(res, e->pos + lo, hi - lo); /* the buckets of hashes [lower, upper), in hash order */
    res->bypass = 1;
    /* tail buckets appended by check_last_kmer (:146): singletons _last_kmer_refs[i] = { n - k + i } */
    for (uint32_t i = 1; i < k - (uint32_t)m + 1; ++i)
        if (pf.tail_match[i]) {
            res->tail[res->n_tail] = (uint32_t)(ix->n - k + i);
            result_push(res, &res->tail[res->n_tail], 1);
            res->n_tail++;
        }
    return KO_OK;
}

/* ---------------------------------------------------------------- choose_search_scheme */
static int cmp_desc_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x < y) - (x > y);
}

/* kmer_index.hpp:407-476 */
static void choose_search_scheme(ko_index *ix)
{
    const uint32_t R = KO_QUERY_SIZE_RANGE;
    memcpy(ix->all_ks, ix->ks, ix->n_ks * sizeof(uint32_t));
    qsort(ix->all_ks, ix->n_ks, sizeof(uint32_t), cmp_desc_u32); /* :410 */
    uint32_t high_ks[64], n_high = 0;
    for (uint32_t i = 0; i < ix->n_ks; ++i)
        if (ix->all_ks[i] >= 9) /* :414 */
            high_ks[n_high++] = ix->all_ks[i];
    /* optimal[q]: list as (length, last summand, previous q) chain; materialised at the end */
    uint32_t *len = (uint32_t *)calloc(R, sizeof(uint32_t));
    uint32_t *last = (uint32_t *)calloc(R, sizeof(uint32_t));
    uint32_t *prev = (uint32_t *)calloc(R, sizeof(uint32_t));
    ix->use_multi = (uint8_t *)calloc(R, 1);
    for (uint32_t i = 0; i < n_high; ++i) { /* :421-425 */
        uint32_t k = high_ks[i];
        if (k < R) {
            len[k] = 1;
            last[k] = k;
            prev[k] = 0;
            ix->use_multi[k] = 1;
        }
    }
    for (uint32_t q = ix->all_ks[0] + 1; q < R; ++q) { /* :427-443 */
        for (uint32_t i = 0; i < n_high; ++i) {
            uint32_t k = high_ks[i];
            if (len[q - k] != 0) {
                /* :434-435 copy optimal[q-k] and push_back(k). Only lists built in this phase are
                   non-empty here, so the chain (prev) is valid. */
                len[q] = len[q - k] + 1;
                last[q] = k;
                prev[q] = q - k;
                ix->use_multi[q] = 1;
                break;
            }
        }
    }
    for (uint32_t q = 0; q < R; ++q) { /* :445-475 */
        if (len[q] != 0)
            continue;
        uint32_t optimal_k = ix->all_ks[0];
        if (q < ix->all_ks[0]) { /* :450-462 (unsigned arithmetic as in the reference) */
            for (uint32_t i = 0; i < ix->n_ks; ++i) {
                uint64_t k = ix->all_ks[i];
                if (q <= k && (k - q < (uint64_t)optimal_k - q))
                    optimal_k = (uint32_t)k;
            }
        } else { /* :463-474, float arithmetic as in the reference */
            for (uint32_t i = 0; i < ix->n_ks; ++i) {
                uint32_t k = ix->all_ks[i];
                if ((ceilf((float)q / (float)k) * (float)k - (float)q) <
                    (ceilf((float)q / (float)optimal_k) * (float)optimal_k - (float)q))
                    optimal_k = k;
            }
        }
        len[q] = 1;
        last[q] = optimal_k;
        prev[q] = 0;
    }
    ix->sum_off = (uint64_t *)malloc((R + 1) * sizeof(uint64_t));
    uint64_t total = 0;
    for (uint32_t q = 0; q < R; ++q) {
        ix->sum_off[q] = total;
        total += len[q];
    }
    ix->sum_off[R] = total;
    ix->sum_ks = (uint8_t *)malloc(total);
    for (uint32_t q = 0; q < R; ++q) {
        uint64_t o = ix->sum_off[q] + len[q];
        uint32_t c = q;
        for (uint32_t j = 0; j < len[q]; ++j) {
            ix->sum_ks[--o] = (uint8_t)last[c];
            c = prev[c];
        }
    }
    free(len);
    free(last);
    free(prev);
}

static const ko_element *element_for_k(const ko_index *ix, uint32_t k)
{
    for (uint32_t i = 0; i < ix->n_ks; ++i)
        if (ix->ks[i] == k)
            return &ix->elems[i];
    return NULL;
}

/* kmer_index.hpp:505-558  kmer_index::search */
static int index_search(const ko_index *ix, const uint8_t *q, uint64_t m, ko_result *res)
{
    if (m > KO_QUERY_SIZE_RANGE) /* :507-509 */
        return KO_THROW_INVALID_ARGUMENT;
    if (m == 0 || m == KO_QUERY_SIZE_RANGE) /* assert :195 / out-of-bounds table read :512 */
        return KO_UNDEFINED;
    const uint8_t *S = ix->sum_ks + ix->sum_off[m];
    const uint64_t s = ix->sum_off[m + 1] - ix->sum_off[m];
    if (!ix->use_multi[m] || ix->n_ks == 1) /* :512-513 */
        return element_search(ix, element_for_k(ix, S[0]), q, m, res);
    /* :516-527  note `last_k = current_k` (not cumulative) */
    ko_bucket *nk = (ko_bucket *)malloc(s * sizeof(ko_bucket));
    uint64_t last_k = 0;
    for (uint64_t i = 0; i < s; ++i) {
        const ko_element *e = element_for_k(ix, S[i]);
        if (!element_at(e, ko_hash(q + last_k, e->k, ix->sigma), &nk[i])) { /* search_k :182-190 */
            free(nk);
            return KO_OK;
        }
        last_k = S[i];
    }
    if (s == 1) { /* :529-530 */
        result_single(res, &nk[0], 1, 1);
        free(nk);
        return KO_OK;
    }
    result_single(res, &nk[0], 1, 0); /* :532 */
    const uint64_t nk_sum_i = 0;      /* :535, never advanced */
    for (uint64_t sp = 0; sp < nk[0].len; ++sp) { /* :536-555 */
        uint64_t previous_pos = nk[0].p[sp];
        int interrupted = 0;
        for (uint64_t j = 1; j < s; ++j) {
            const ko_bucket *current = &nk[j];
            uint64_t it = bucket_lower_bound(current, previous_pos += S[nk_sum_i]);
            if (it == current->len)
                ko_ub_seen = 1;
            if (it == current->len || (uint64_t)current->p[it] != previous_pos) { /* :546 */
                interrupted = 1;
                break;
            }
        }
        if (interrupted)
            res->mask[sp] = 0;
    }
    free(nk);
    return KO_OK;
}

/* ---------------------------------------------------------------- public C API */
ko_index *ko_create(const uint8_t *ranks, uint64_t n, uint32_t sigma, const uint32_t *ks, uint32_t n_ks)
{
    if (n_ks == 0 || n_ks > 64)
        return NULL;
    for (uint32_t i = 0; i < n_ks; ++i)
        if (ks[i] == 0 || ks[i] > 63 || n < ks[i] || !((double)ks[i] < 64.0 / log2((double)sigma))) /* :42 */
            return NULL;
    ko_index *ix = (ko_index *)calloc(1, sizeof(ko_index));
    ix->sigma = sigma;
    ix->n = n;
    ix->text = (uint8_t *)malloc(n);
    memcpy(ix->text, ranks, n);
    ix->n_ks = n_ks;
    memcpy(ix->ks, ks, n_ks * sizeof(uint32_t));
    ix->elems = (ko_element *)calloc(n_ks, sizeof(ko_element));
    for (uint32_t i = 0; i < n_ks; ++i)
        if (element_create(&ix->elems[i], ks[i], ix->text, n, sigma) != 0)
            return NULL;
    choose_search_scheme(ix);
    return ix;
}

/* ---------------------------------------------------------------- restricted index (still TEST INFRASTRUCTURE)
   The same index, holding only the buckets a given query batch can ask for: `_data` restricted to the hashes that
   index_search() / element_search() pass to at() (kmer_index.hpp:76-84) for these queries, plus the hash ranges of
   their prefix enumerations (:131-144). Every bucket held is complete (all its positions, ascending), so every
   lookup the batch makes returns exactly what the full index returns; a lookup of a bucket that is not held is
   counted in ko_restricted_misses() (tests assert 0). This is what lets a CPU check a 10^6-query sample against
   the 3 Gbp text of BASELINE config 5: one threaded scan of the text instead of a 36 GB sorted pair array.
   tests/test_oracle_golden.py pins restricted == full on every golden fixture. */
typedef struct {
    const uint8_t *text;
    uint64_t lo, hi; /* k-mer start positions [lo, hi) */
    uint32_t k, sigma;
    const uint64_t *needed;
    uint64_t *hashes; /* thread-local output */
    uint32_t *pos;
    uint64_t n, cap;
} ko_scan_job;

static void *scan_worker(void *arg)
{
    ko_scan_job *j = (ko_scan_job *)arg;
    if (j->lo >= j->hi)
        return NULL;
    const uint64_t top = ko_fast_pow(j->sigma, (uint8_t)(j->k - 1));
    uint64_t h = ko_hash(j->text + j->lo, j->k, j->sigma);
    for (uint64_t p = j->lo;; ++p) {
        if ((j->needed[h >> 6] >> (h & 63)) & 1) {
            if (j->n == j->cap) {
                j->cap = j->cap ? 2 * j->cap : 4096;
                j->hashes = (uint64_t *)realloc(j->hashes, j->cap * sizeof(uint64_t));
                j->pos = (uint32_t *)realloc(j->pos, j->cap * sizeof(uint32_t));
            }
            j->hashes[j->n] = h;
            j->pos[j->n++] = (uint32_t)p;
        }
        if (p + 1 >= j->hi)
            break;
        h = (h - (uint64_t)j->text[p] * top) * j->sigma + j->text[p + j->k];
    }
    return NULL;
}

static void mark_needed(ko_element *e, uint64_t key_space, uint64_t lo, uint64_t hi)
{
    if (hi > key_space)
        hi = key_space;
    for (uint64_t h = lo; h < hi; ++h)
        e->needed[h >> 6] |= 1ull << (h & 63);
}

ko_index *ko_create_restricted(const uint8_t *ranks, uint64_t n, uint32_t sigma, const uint32_t *ks, uint32_t n_ks,
                               const uint8_t *q, const uint64_t *q_off, uint64_t Q, uint32_t n_threads)
{
    if (n_ks == 0 || n_ks > 64)
        return NULL;
    for (uint32_t i = 0; i < n_ks; ++i) {
        if (ks[i] == 0 || ks[i] > 63 || n < ks[i] || !((double)ks[i] < 64.0 / log2((double)sigma)))
            return NULL;
        if (pow((double)sigma, (double)ks[i]) > 68719476736.0) /* the bucket bitmap: at most 2^36 bits = 8 GB */
            return NULL;
    }
    if (n_threads == 0)
        n_threads = 1;
    if (n_threads > 256)
        n_threads = 256;
    ko_index *ix = (ko_index *)calloc(1, sizeof(ko_index));
    ix->sigma = sigma;
    ix->n = n;
    ix->text = (uint8_t *)malloc(n);
    memcpy(ix->text, ranks, n);
    ix->n_ks = n_ks;
    memcpy(ix->ks, ks, n_ks * sizeof(uint32_t));
    ix->elems = (ko_element *)calloc(n_ks, sizeof(ko_element));
    choose_search_scheme(ix);
    uint64_t key_space[64];
    for (uint32_t i = 0; i < n_ks; ++i) {
        ko_element *e = &ix->elems[i];
        e->k = ks[i];
        key_space[i] = ko_fast_pow(sigma, (uint8_t)ks[i]);
        e->needed = (uint64_t *)calloc(key_space[i] / 64 + 2, sizeof(uint64_t));
        e->last_kmer = (uint8_t *)malloc(ks[i]);
        memcpy(e->last_kmer, ix->text + n - ks[i], ks[i]);
        if (!e->needed)
            return NULL;
    }
    /* the lookups of index_search (kmer_index.hpp:505-558) / element_search (:193-346), without their early returns */
    for (uint64_t i = 0; i < Q; ++i) {
        const uint8_t *qq = q + q_off[i];
        const uint64_t m = q_off[i + 1] - q_off[i];
        if (m == 0 || m >= KO_QUERY_SIZE_RANGE)
            continue;
        const uint8_t *S = ix->sum_ks + ix->sum_off[m];
        const uint64_t s = ix->sum_off[m + 1] - ix->sum_off[m];
        if (!ix->use_multi[m] || n_ks == 1) {
            uint32_t ei = 0;
            for (uint32_t c = 0; c < n_ks; ++c)
                if (ks[c] == S[0])
                    ei = c;
            ko_element *e = &ix->elems[ei];
            const uint64_t k = e->k;
            if (m >= k) { /* :198-205, :216-227: at(hash) of every full part */
                for (uint64_t j = 0; j < m / k; ++j) {
                    const uint64_t h = ko_hash(qq + j * k, (uint32_t)k, sigma);
                    mark_needed(e, key_space[ei], h, h + 1);
                }
            } else if (!((double)ko_fast_pow(sigma, (uint8_t)(k - m)) > 1e7)) { /* :342-345 -> :131-144 */
                uint64_t lower = 0;
                for (uint64_t c = 0; c < m; ++c)
                    lower += (uint64_t)qq[c] * ko_fast_pow(sigma, (uint8_t)(k - c - 1));
                mark_needed(e, key_space[ei], lower, lower + ko_fast_pow(sigma, (uint8_t)(k - m)));
            }
        } else { /* :516-527, `last_k = current_k` */
            uint64_t last_k = 0;
            for (uint64_t c = 0; c < s; ++c) {
                uint32_t ei = 0;
                for (uint32_t d = 0; d < n_ks; ++d)
                    if (ks[d] == S[c])
                        ei = d;
                if (last_k + S[c] <= m) {
                    const uint64_t h = ko_hash(qq + last_k, S[c], sigma);
                    mark_needed(&ix->elems[ei], key_space[ei], h, h + 1);
                }
                last_k = S[c];
            }
        }
    }
    /* create (:154-179) restricted to the needed buckets: threaded scan, chunks concatenated in text order */
    for (uint32_t i = 0; i < n_ks; ++i) {
        ko_element *e = &ix->elems[i];
        const uint64_t n_kmers = n - e->k + 1;
        ko_scan_job jobs[256];
        pthread_t th[256];
        for (uint32_t t = 0; t < n_threads; ++t) {
            memset(&jobs[t], 0, sizeof(ko_scan_job));
            jobs[t].text = ix->text;
            jobs[t].lo = n_kmers * t / n_threads;
            jobs[t].hi = n_kmers * (t + 1) / n_threads;
            jobs[t].k = e->k;
            jobs[t].sigma = sigma;
            jobs[t].needed = e->needed;
            pthread_create(&th[t], NULL, scan_worker, &jobs[t]);
        }
        uint64_t kept = 0;
        for (uint32_t t = 0; t < n_threads; ++t) {
            pthread_join(th[t], NULL);
            kept += jobs[t].n;
        }
        e->n_kmers = kept;
        e->hashes = (uint64_t *)malloc((kept ? kept : 1) * sizeof(uint64_t));
        e->pos = (uint32_t *)malloc((kept ? kept : 1) * sizeof(uint32_t));
        uint64_t o = 0, max_key = 0;
        for (uint32_t t = 0; t < n_threads; ++t) {
            if (jobs[t].n) {
                memcpy(e->hashes + o, jobs[t].hashes, jobs[t].n * sizeof(uint64_t));
                memcpy(e->pos + o, jobs[t].pos, jobs[t].n * sizeof(uint32_t));
            }
            o += jobs[t].n;
            free(jobs[t].hashes);
            free(jobs[t].pos);
        }
        for (uint64_t c = 0; c < kept; ++c)
            if (e->hashes[c] > max_key)
                max_key = e->hashes[c];
        sort_pairs(e->hashes, e->pos, kept, max_key);
    }
    return ix;
}

uint64_t ko_restricted_misses(void)
{
    return ko_restricted_misses_;
}

void ko_destroy(ko_index *ix)
{
    if (!ix)
        return;
    for (uint32_t i = 0; i < ix->n_ks; ++i) {
        free(ix->elems[i].hashes);
        free(ix->elems[i].pos);
        free(ix->elems[i].last_kmer);
        free(ix->elems[i].needed);
    }
    free(ix->elems);
    free(ix->text);
    free(ix->sum_off);
    free(ix->sum_ks);
    free(ix->use_multi);
    free(ix);
}

uint64_t ko_scheme(const ko_index *ix, uint64_t m, uint32_t *out, uint64_t cap, int *use_multi)
{
    uint64_t s = ix->sum_off[m + 1] - ix->sum_off[m];
    for (uint64_t i = 0; i < s && i < cap; ++i)
        out[i] = ix->sum_ks[ix->sum_off[m] + i];
    *use_multi = ix->use_multi[m];
    return s;
}

typedef struct {
    const ko_index *ix;
    const uint8_t *q;
    const uint64_t *q_off;
    uint64_t lo, hi;
    uint64_t *counts;
    uint8_t *status;
    uint32_t *pos;
    uint64_t n_pos, cap_pos;
    int keep;
} ko_job;

static void *search_worker(void *arg)
{
    ko_job *j = (ko_job *)arg;
    for (uint64_t i = j->lo; i < j->hi; ++i) {
        ko_result r;
        result_init(&r);
        ko_ub_seen = 0;
        int st = index_search(j->ix, j->q + j->q_off[i], j->q_off[i + 1] - j->q_off[i], &r);
        j->status[i] = (uint8_t)(st | (ko_ub_seen ? KO_UB_FLAG : 0));
        uint32_t *v = NULL;
        uint64_t c = 0;
        if (st == KO_OK && !j->keep && r.bypass) {
            /* count only: a bypass result (exact-k, sub-k) is every element of its buckets (kmer_index_result.hpp:250) */
            for (uint64_t b = 0; b < r.n_buckets; ++b)
                c += r.buckets[b].len;
        } else if (st == KO_OK)
            c = result_to_vector(&r, &v);
        j->counts[i] = c;
        if (j->keep && c) {
            if (j->n_pos + c > j->cap_pos) {
                j->cap_pos = 2 * (j->n_pos + c);
                j->pos = (uint32_t *)realloc(j->pos, j->cap_pos * sizeof(uint32_t));
            }
            memcpy(j->pos + j->n_pos, v, c * sizeof(uint32_t));
            j->n_pos += c;
        }
        free(v);
        result_free(&r);
    }
    return NULL;
}

/* same contract as kref_search_batch in oracle/ref_driver.cpp */
int ko_search_batch(const ko_index *ix, const uint8_t *q, const uint64_t *q_off, uint64_t Q, uint32_t n_threads,
                    uint64_t *counts, uint8_t *status, uint32_t **positions, uint64_t *total, int keep_positions)
{
    if (n_threads == 0)
        n_threads = 1;
    if (n_threads > 256)
        n_threads = 256;
    ko_job jobs[256];
    pthread_t th[256];
    for (uint32_t t = 0; t < n_threads; ++t) {
        memset(&jobs[t], 0, sizeof(ko_job));
        jobs[t].ix = ix;
        jobs[t].q = q;
        jobs[t].q_off = q_off;
        jobs[t].lo = Q * t / n_threads;
        jobs[t].hi = Q * (t + 1) / n_threads;
        jobs[t].counts = counts;
        jobs[t].status = status;
        jobs[t].keep = keep_positions;
        pthread_create(&th[t], NULL, search_worker, &jobs[t]);
    }
    uint64_t tot = 0;
    for (uint32_t t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        tot += jobs[t].n_pos;
    }
    if (total)
        *total = tot;
    if (positions) {
        *positions = (uint32_t *)malloc((tot ? tot : 1) * sizeof(uint32_t));
        uint64_t o = 0;
        for (uint32_t t = 0; t < n_threads; ++t) {
            if (jobs[t].n_pos)
                memcpy(*positions + o, jobs[t].pos, jobs[t].n_pos * sizeof(uint32_t));
            o += jobs[t].n_pos;
        }
    }
    for (uint32_t t = 0; t < n_threads; ++t)
        free(jobs[t].pos);
    return 0;
}

void ko_free(void *p)
{
    free(p);
}

/* the index content itself, for CSR parity: positions stably sorted by hash and the sorted hashes */
uint64_t ko_element_size(const ko_index *ix, uint32_t elem)
{
    return ix->elems[elem].n_kmers;
}
const uint32_t *ko_element_positions(const ko_index *ix, uint32_t elem)
{
    return ix->elems[elem].pos;
}
const uint64_t *ko_element_hashes(const ko_index *ix, uint32_t elem)
{
    return ix->elems[elem].hashes;
}

/* ---------------------------------------------------------------- ground truth (NOT the reference):
   occ(q) = sorted { p : T[p, p+m) == q }, by plain scan. Used to document where the reference is wrong
   and to gate the product's optional CORRECT mode. */
int ko_truth_search_batch(const uint8_t *text, uint64_t n, const uint8_t *q, const uint64_t *q_off, uint64_t Q,
                          uint64_t *counts, uint32_t **positions, uint64_t *total)
{
    uint64_t cap = 1024, np = 0;
    uint32_t *pos = (uint32_t *)malloc(cap * sizeof(uint32_t));
    for (uint64_t i = 0; i < Q; ++i) {
        const uint8_t *qq = q + q_off[i];
        uint64_t m = q_off[i + 1] - q_off[i];
        uint64_t c = 0;
        if (m > 0 && m <= n) {
            const uint8_t *p = text, *end = text + n - m + 1;
            while (p < end && (p = (const uint8_t *)memchr(p, qq[0], (size_t)(end - p))) != NULL) {
                if (memcmp(p, qq, m) == 0) {
                    if (np == cap) {
                        cap *= 2;
                        pos = (uint32_t *)realloc(pos, cap * sizeof(uint32_t));
                    }
                    pos[np++] = (uint32_t)(p - text);
                    ++c;
                }
                ++p;
            }
        }
        counts[i] = c;
    }
    *positions = pos;
    *total = np;
    return 0;
}

/* ---------------------------------------------------------------- choose_best_k.hpp:12-60
   The reference sorts with unstable std::sort on the score only (:50-51); ties are broken here by the
   candidate order {29,...,10} (stable), which is what libstdc++'s insertion sort does for 10 elements. */
uint64_t ko_choose_best_k(const uint64_t *lens, uint64_t n_lens, uint64_t n_k, uint64_t *out)
{
    static const uint64_t cand[10] = {29, 27, 25, 23, 21, 19, 17, 13, 11, 10}; /* :23 */
    uint64_t score[10] = {0};
    for (uint64_t a = 0; a < n_lens; ++a) {
        uint64_t i = lens[a];
        for (int c = 0; c < 10; ++c) {
            uint64_t k = cand[c];
            if (i % k == 0) { /* :33-37 */
                score[c] += 3;
                break;
            } else if (k - (i % k) <= 3) { /* :39-43 */
                score[c] += 4 - (k - (i % k));
                break;
            }
        }
    }
    int order[10];
    for (int c = 0; c < 10; ++c)
        order[c] = c;
    for (int a = 1; a < 10; ++a) { /* stable insertion sort, descending score */
        int o = order[a], b = a;
        while (b > 0 && score[order[b - 1]] < score[o]) {
            order[b] = order[b - 1];
            --b;
        }
        order[b] = o;
    }
    uint64_t w = 0;
    for (; w < n_k && w < 10; ++w)
        out[w] = cand[order[w]];
    return w;
}
