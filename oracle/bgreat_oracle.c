/* TEST INFRASTRUCTURE ONLY -- see bgreat_oracle.h. Parity status: PINNED against the
 * compiled reference (tests/golden/, tests/test_oracle_vs_ref.py).
 *
 * Plain-C restatement of the reference's read-to-graph mapping path, string based like
 * the reference. The reference's BooPHF MPHF is replaced by a sorted key array + binary
 * search: every reference lookup is verified against the stored key (aligner.cpp:158,166,
 * 219,227,353,361), so any exact dictionary yields the same results.
 *
 * Reference files restated (all under /root/reference):
 *   utils.cpp:52-73,117-192   sequence primitives
 *   aligner.cpp:46-117        getReads
 *   aligner.cpp:147-267       getEnd / getBegin
 *   aligner.cpp:305-378       update / updateRC / getListOverlap / getNOverlap
 *   aligner.cpp:407-534       indexUnitigsAux
 *   aligner.cpp:550-609       alignAll / printPath
 *   alignerGreedy.cpp:35-57,167-431
 *   alignerExhaustive.cpp:35-318
 *   alignerPaths.cpp:35-460
 */
#include "bgreat_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small containers */

typedef struct {
    int32_t* v;
    int n, cap;
} ivec;

static void iv_push(ivec* p, int32_t x) {
    if (p->n == p->cap) {
        p->cap = p->cap ? 2 * p->cap : 16;
        p->v = (int32_t*)realloc(p->v, sizeof(int32_t) * (size_t)p->cap);
    }
    p->v[p->n++] = x;
}
static void iv_append(ivec* p, const ivec* q) {
    for (int i = 0; i < q->n; ++i) iv_push(p, q->v[i]);
}
static void iv_assign(ivec* p, const ivec* q) {
    p->n = 0;
    iv_append(p, q);
}
static void iv_free(ivec* p) {
    free(p->v);
    p->v = NULL;
    p->n = p->cap = 0;
}
static int iv_contains(const ivec* p, int32_t x) {
    for (int i = 0; i < p->n; ++i)
        if (p->v[i] == x) return 1;
    return 0;
}
static void iv_reverse(ivec* p) {
    for (int i = 0, j = p->n - 1; i < j; ++i, --j) {
        int32_t t = p->v[i];
        p->v[i] = p->v[j];
        p->v[j] = t;
    }
}

typedef struct {
    char* s;
    size_t n, cap;
} sbuf;

static void sb_reserve(sbuf* b, size_t need) {
    if (need > b->cap) {
        size_t c = b->cap ? b->cap : 64;
        while (c < need) c *= 2;
        b->s = (char*)realloc(b->s, c);
        b->cap = c;
    }
}
static void sb_putc(sbuf* b, char c) {
    sb_reserve(b, b->n + 1);
    b->s[b->n++] = c;
}
static void sb_put(sbuf* b, const char* s, size_t n) {
    sb_reserve(b, b->n + n);
    memcpy(b->s + b->n, s, n);
    b->n += n;
}

/* ------------------------------------------------------------------ sequence primitives */

/* str2num, utils.cpp:117-129: A0 C1 G2, anything else 3; first base most significant. */
uint64_t bo_str2num(const char* s, uint32_t n) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < n; ++i) {
        r <<= 2;
        if (s[i] == 'A') r += 0;
        else if (s[i] == 'C') r += 1;
        else if (s[i] == 'G') r += 2;
        else r += 3;
    }
    return r;
}

/* rcb, utils.cpp:182-192: reverse complement of an n-base 2-bit code. */
uint64_t bo_rcb(uint64_t x, uint32_t n) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < n; ++i) {
        r = (r << 2) | (3 - (x & 3));
        x >>= 2;
    }
    return r;
}

/* nuc2int / nuc2intrc, utils.cpp:132-151: the rolling encoders send non-CGT (resp. non-ACG) to 0. */
static uint64_t nuc_fwd(char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }
static uint64_t nuc_rc(char c) { return c == 'A' ? 3 : c == 'C' ? 2 : c == 'G' ? 1 : 0; }

/* reverseComplements / revCompChar, utils.cpp:52-73: non-ACG -> 'A' (so T->A and N->A). */
void bo_reverse_complement(const char* s, uint32_t n, char* out) {
    for (uint32_t i = 0; i < n; ++i) {
        char c = s[n - 1 - i];
        out[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
    }
}

/* missmatchNumber, utils.cpp:154-168: mismatches over |seq2| characters, leaving as soon as
 * the count exceeds n (the value returned is then n+1). */
uint32_t bo_mismatch_number(const char* seq1, const char* seq2, uint32_t n2, uint32_t n) {
    uint32_t miss = 0;
    for (uint32_t i = 0; i < n2; ++i) {
        if (seq2[i] != seq1[i]) {
            if (++miss > n) return miss;
        }
    }
    return miss;
}

/* ------------------------------------------------------------------ aligner state */

typedef struct {
    uint64_t overlap; /* unitigIndices, aligner.h:49-55 */
    uint32_t id[4];
} slots_t;

typedef struct {
    uint64_t* keys; /* sorted, unique: stands in for the MPHF */
    slots_t* ent;
    uint64_t n;
} table_t;

struct bo_aligner {
    uint32_t k, errors, effort;
    char** unitigs; /* [0] is the "" sentinel: ids are 1-based (aligner.cpp:408) */
    uint32_t* ulen;
    uint32_t n_unitigs; /* including the sentinel */
    uint32_t cap_unitigs;
    table_t left, right;
    uint64_t c[BO_C_COUNT];
};

#define NOT_FOUND UINT64_MAX

static uint64_t tbl_lookup(bo_aligner* a, const table_t* t, uint64_t key) {
    a->c[BO_C_LOOKUPS]++;
    uint64_t lo = 0, hi = t->n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (t->keys[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return (lo < t->n && t->keys[lo] == key) ? lo : NOT_FOUND;
}

bo_aligner* bo_new(uint32_t k, uint32_t errors, uint32_t effort) {
    bo_aligner* a = (bo_aligner*)calloc(1, sizeof(bo_aligner));
    a->k = k;
    a->errors = errors;
    a->effort = effort;
    return a;
}

/* The index does not depend on -m / -e (aligner.cpp:407-534 never looks at them): the same indexed graph can be
 * asked with other settings (test convenience: one index, a sweep over -m). */
void bo_set_params(bo_aligner* a, uint32_t errors, uint32_t effort) {
    a->errors = errors;
    a->effort = effort;
}

void bo_free(bo_aligner* a) {
    if (!a) return;
    for (uint32_t i = 0; i < a->n_unitigs; ++i) free(a->unitigs[i]);
    free(a->unitigs);
    free(a->ulen);
    free(a->left.keys);
    free(a->left.ent);
    free(a->right.keys);
    free(a->right.ent);
    free(a);
}

void bo_counters(const bo_aligner* a, uint64_t out[BO_C_COUNT]) { memcpy(out, a->c, sizeof(a->c)); }
void bo_reset_counters(bo_aligner* a) { memset(a->c, 0, sizeof(a->c)); }

/* ------------------------------------------------------------------ istream emulation
 * The reference parses with std::ifstream + std::getline + peek; its behaviour at end of
 * file (phantom FASTQ read, no phantom FASTA read) follows from the iostream state bits,
 * so they are modelled exactly:
 *   getline on a stream that is not good(): sets failbit, leaves the string untouched;
 *   getline on a good stream: erases the string, extracts up to and including '\n';
 *     hitting end of input sets eofbit; extracting nothing at all also sets failbit;
 *   peek on a good stream at end of input sets eofbit and yields EOF. */

typedef struct {
    const char* d;
    size_t n, pos;
    int eofbit, failbit;
} istream_t;

static int is_good(const istream_t* s) { return !s->eofbit && !s->failbit; }

static void is_getline(istream_t* s, sbuf* out) {
    if (!is_good(s)) {
        s->failbit = 1;
        return;
    }
    out->n = 0;
    size_t extracted = 0;
    for (;;) {
        if (s->pos >= s->n) {
            s->eofbit = 1;
            break;
        }
        char c = s->d[s->pos++];
        ++extracted;
        if (c == '\n') break;
        sb_putc(out, c);
    }
    if (extracted == 0) s->failbit = 1;
}

static int is_peek(istream_t* s) {
    if (!is_good(s)) return -1;
    if (s->pos >= s->n) {
        s->eofbit = 1;
        return -1;
    }
    return (unsigned char)s->d[s->pos];
}

static char* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    sbuf b = {0, 0, 0};
    char tmp[1 << 16];
    size_t r;
    while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) sb_put(&b, tmp, r);
    fclose(f);
    sb_reserve(&b, b.n + 1);
    *n = b.n;
    return b.s;
}

/* ------------------------------------------------------------------ index build
 * indexUnitigsAux, aligner.cpp:407-534. */

static int cmp_u64(const void* x, const void* y) {
    uint64_t a = *(const uint64_t*)x, b = *(const uint64_t*)y;
    return a < b ? -1 : a > b;
}

static void push_unitig(bo_aligner* a, const char* s, uint32_t n) {
    if (a->n_unitigs == a->cap_unitigs) {
        a->cap_unitigs = a->cap_unitigs ? 2 * a->cap_unitigs : 1024;
        a->unitigs = (char**)realloc(a->unitigs, sizeof(char*) * a->cap_unitigs);
        a->ulen = (uint32_t*)realloc(a->ulen, sizeof(uint32_t) * a->cap_unitigs);
    }
    char* c = (char*)malloc((size_t)n + 1);
    memcpy(c, s, n);
    c[n] = 0;
    a->unitigs[a->n_unitigs] = c;
    a->ulen[a->n_unitigs] = n;
    a->n_unitigs++;
}

static void tbl_from_keys(table_t* t, uint64_t* keys, uint64_t n) {
    /* sort + unique, aligner.cpp:445-448 */
    qsort(keys, n, sizeof(uint64_t), cmp_u64);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (m == 0 || keys[m - 1] != keys[i]) keys[m++] = keys[i];
    t->keys = keys;
    t->n = m;
    t->ent = (slots_t*)calloc(m ? m : 1, sizeof(slots_t)); /* {0,0,0,0,0}, aligner.cpp:463-464 */
}

/* first free slot, the 5th and later arrivals overwrite slot 4 (aligner.cpp:481-489) */
static void slot_insert(slots_t* e, uint64_t key, uint32_t id) {
    e->overlap = key;
    if (e->id[0] == 0) e->id[0] = id;
    else if (e->id[1] == 0) e->id[1] = id;
    else if (e->id[2] == 0) e->id[2] = id;
    else e->id[3] = id;
}

static void build_tables(bo_aligner* a) {
    const uint32_t k = a->k;
    uint64_t nu = a->n_unitigs - 1;
    uint64_t* lk = (uint64_t*)malloc(sizeof(uint64_t) * (2 * nu + 1));
    uint64_t* rk = (uint64_t*)malloc(sizeof(uint64_t) * (2 * nu + 1));
    uint64_t nl = 0, nr = 0;
    /* first pass: key sets (aligner.cpp:422-433). A begin that is canonical goes left, otherwise
     * its reverse complement goes right; an end that is canonical goes right, otherwise its
     * reverse complement goes left. Ties (palindromes) are "canonical". */
    for (uint32_t i = 1; i < a->n_unitigs; ++i) {
        const char* u = a->unitigs[i];
        uint32_t n = a->ulen[i];
        uint64_t beg = bo_str2num(u, k - 1), rcbeg = bo_rcb(beg, k - 1);
        if (beg <= rcbeg) lk[nl++] = beg;
        else rk[nr++] = rcbeg;
        uint64_t end = bo_str2num(u + n - (k - 1), k - 1), rcend = bo_rcb(end, k - 1);
        if (end <= rcend) rk[nr++] = end;
        else lk[nl++] = rcend;
    }
    tbl_from_keys(&a->left, lk, nl);
    tbl_from_keys(&a->right, rk, nr);
    /* second pass in file order: slot fill (aligner.cpp:466-533) */
    for (uint32_t i = 1; i < a->n_unitigs; ++i) {
        const char* u = a->unitigs[i];
        uint32_t n = a->ulen[i];
        uint64_t beg = bo_str2num(u, k - 1), rcbeg = bo_rcb(beg, k - 1);
        if (beg <= rcbeg) slot_insert(&a->left.ent[tbl_lookup(a, &a->left, beg)], beg, i);
        else slot_insert(&a->right.ent[tbl_lookup(a, &a->right, rcbeg)], rcbeg, i);
        uint64_t end = bo_str2num(u + n - (k - 1), k - 1), rcend = bo_rcb(end, k - 1);
        if (end <= rcend) slot_insert(&a->right.ent[tbl_lookup(a, &a->right, end)], end, i);
        else slot_insert(&a->left.ent[tbl_lookup(a, &a->left, rcend)], rcend, i);
    }
    a->c[BO_C_LOOKUPS] = 0;
}

int64_t bo_index_file(bo_aligner* a, const char* unitig_path) {
    push_unitig(a, "", 0);
    size_t n = 0;
    char* data = slurp(unitig_path, &n);
    istream_t s = {data, n, 0, 0, 0};
    if (!data) s.failbit = 1; /* open failed: every getline fails, line stays "" */
    sbuf line = {0, 0, 0};
    /* aligner.cpp:415-421: two getlines per record, only the second one is kept; the loop
     * ends at end of file or at the first kept line shorter than k. */
    while (!s.eofbit) {
        is_getline(&s, &line);
        is_getline(&s, &line);
        if (line.n < a->k) break;
        push_unitig(a, line.s, (uint32_t)line.n);
    }
    free(line.s);
    free(data);
    build_tables(a);
    return data ? (int64_t)a->n_unitigs - 1 : -1;
}

int64_t bo_index_mem(bo_aligner* a, const char* seqs, const uint64_t* off, uint32_t n) {
    push_unitig(a, "", 0);
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = off[i + 1] - off[i];
        if (len < a->k) break;
        push_unitig(a, seqs + off[i], (uint32_t)len);
    }
    build_tables(a);
    return (int64_t)a->n_unitigs - 1;
}

/* ------------------------------------------------------------------ index queries
 * getEnd / getBegin, aligner.cpp:147-267: oriented unitigs that end / begin with the
 * (k-1)-mer `bin`. Candidates are copies (forward, or reverse-complemented with a negative
 * id); the orientation is decided by the forward test only; slots are scanned in order up
 * to the first empty one. */

typedef struct {
    char* seq;
    uint32_t len;
    int32_t id;
} cand_t;

typedef struct {
    cand_t c[4];
    int n;
} cands_t;

static void cands_free(cands_t* cl) {
    for (int i = 0; i < cl->n; ++i) free(cl->c[i].seq);
    cl->n = 0;
}

static void cands_add(bo_aligner* a, cands_t* cl, uint32_t id, int forward) {
    uint32_t n = a->ulen[id];
    char* s = (char*)malloc((size_t)n + 1);
    if (forward) memcpy(s, a->unitigs[id], n);
    else bo_reverse_complement(a->unitigs[id], n, s);
    s[n] = 0;
    cl->c[cl->n].seq = s;
    cl->c[cl->n].len = n;
    cl->c[cl->n].id = forward ? (int32_t)id : -(int32_t)id;
    cl->n++;
}

static void query(bo_aligner* a, uint64_t bin, int want_end, cands_t* cl) {
    const uint32_t k = a->k;
    cl->n = 0;
    a->c[BO_C_QUERIES]++;
    uint64_t rc = bo_rcb(bin, k - 1);
    const slots_t* e = NULL;
    /* getEnd: canonical bin -> right table, else left table under rc (aligner.cpp:154-170);
     * getBegin: canonical bin -> left table, else right table under rc (aligner.cpp:215-231). */
    const table_t* canon_tbl = want_end ? &a->right : &a->left;
    const table_t* other_tbl = want_end ? &a->left : &a->right;
    if (bin <= rc) {
        uint64_t h = tbl_lookup(a, canon_tbl, bin);
        if (h != NOT_FOUND && canon_tbl->ent[h].overlap == bin) e = &canon_tbl->ent[h];
    } else {
        uint64_t h = tbl_lookup(a, other_tbl, rc);
        if (h != NOT_FOUND && other_tbl->ent[h].overlap == rc) e = &other_tbl->ent[h];
    }
    if (!e) return;
    for (int s = 0; s < 4; ++s) {
        uint32_t id = e->id[s];
        if (id == 0) break;
        const char* u = a->unitigs[id];
        uint32_t n = a->ulen[id];
        uint64_t code = want_end ? bo_str2num(u + n - (k - 1), k - 1) : bo_str2num(u, k - 1);
        cands_add(a, cl, id, code == bin);
    }
}

static uint32_t mismatches(bo_aligner* a, const char* s1, const char* s2, uint32_t n2, uint32_t n) {
    a->c[BO_C_COMPARES]++;
    a->c[BO_C_CMPBASES] += n2;
    return bo_mismatch_number(s1, s2, n2, n);
}

/* ------------------------------------------------------------------ seed scans */

typedef struct {
    uint64_t code; /* forward code of the window (not the canonical one) */
    uint32_t pos;
} seed_t;

typedef struct {
    seed_t* v;
    uint32_t n, cap;
} seeds_t;

static void seeds_push(seeds_t* s, uint64_t code, uint32_t pos) {
    if (s->n == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 64;
        s->v = (seed_t*)realloc(s->v, sizeof(seed_t) * s->cap);
    }
    s->v[s->n].code = code;
    s->v[s->n].pos = pos;
    s->n++;
}

/* substr(0,k-1) of a possibly shorter read */
static uint32_t min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* update / updateRC, aligner.cpp:305-315 */
static void roll(const bo_aligner* a, uint64_t* num, uint64_t* rcnum, char c) {
    const uint32_t k = a->k;
    uint64_t modulus_bits = 2 * (uint64_t)(k - 1);
    *num = (*num << 2) + nuc_fwd(c);
    if (modulus_bits < 64) *num &= (((uint64_t)1) << modulus_bits) - 1; /* % offsetUpdate */
    *rcnum = (*rcnum >> 2) + (nuc_rc(c) << (2 * k - 4));
}

/* getNOverlap, aligner.cpp:345-378: the first n windows whose canonical (k-1)-mer is a key of
 * the left table or else of the right table. */
static void scan_n_overlap(bo_aligner* a, const char* read, uint32_t len, uint32_t n, seeds_t* out) {
    const uint32_t k = a->k;
    out->n = 0;
    uint64_t num = bo_str2num(read, min_u32(k - 1, len));
    uint64_t rcnum = bo_rcb(num, k - 1);
    uint64_t rep = num < rcnum ? num : rcnum;
    for (uint32_t i = 0;; ++i) {
        a->c[BO_C_WINDOWS]++;
        int done = 0;
        uint64_t h = tbl_lookup(a, &a->left, rep);
        if (h != NOT_FOUND && a->left.ent[h].overlap == rep) {
            seeds_push(out, num, i);
            done = 1;
        }
        if (!done) {
            h = tbl_lookup(a, &a->right, rep);
            if (h != NOT_FOUND && a->right.ent[h].overlap == rep) seeds_push(out, num, i);
        }
        if (out->n >= n) return;
        if ((uint64_t)i + k - 1 < len) {
            roll(a, &num, &rcnum, read[i + k - 1]);
            rep = num < rcnum ? num : rcnum;
        } else {
            return;
        }
    }
}

/* getListOverlap, aligner.cpp:318-342: every window, unfiltered (one unused lookup each). */
static void scan_all(bo_aligner* a, const char* read, uint32_t len, seeds_t* out) {
    const uint32_t k = a->k;
    out->n = 0;
    uint64_t num = bo_str2num(read, min_u32(k - 1, len));
    uint64_t rcnum = bo_rcb(num, k - 1);
    for (uint32_t i = 0;; ++i) {
        a->c[BO_C_WINDOWS]++;
        a->c[BO_C_LOOKUPS]++; /* leftMPHF.lookup(rep), result unused (aligner.cpp:324) */
        seeds_push(out, num, i);
        if ((uint64_t)i + k - 1 < len) roll(a, &num, &rcnum, read[i + k - 1]);
        else return;
    }
}

/* ------------------------------------------------------------------ greedy extension */

/* checkBeginGreedy (alignerGreedy.cpp:268-319) and mapOnLeftEndGreedy (:167-218) are the same
 * procedure: walk leftwards from the anchor (code, pos). */
static uint32_t greedy_left(bo_aligner* a, const char* read, uint64_t code, uint32_t pos, ivec* path,
                            uint32_t errors) {
    const uint32_t k = a->k;
    if (pos == 0) {
        iv_push(path, 0);
        return 0;
    }
    const char* read_left = read; /* read[0, pos) */
    cands_t cl;
    query(a, code, 1, &cl);
    uint32_t best = errors + 1, best_i = 9, next_len = 0;
    int ended = 0;
    int32_t offset = 0;
    uint64_t next_code = 0;
    for (int i = 0; i < cl.n; ++i) {
        const char* u = cl.c[i].seq;
        uint32_t L = cl.c[i].len - (k - 1); /* bases of the unitig left of the overlap */
        if (L >= pos) {
            /* the unitig covers the whole left part of the read */
            uint32_t miss = mismatches(a, u + (L - pos), read_left, pos, errors);
            if (miss == 0) {
                iv_push(path, cl.c[i].id);
                iv_push(path, (int32_t)(L - pos));
                cands_free(&cl);
                return 0;
            } else if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                ended = 1;
                offset = (int32_t)(L - pos);
            }
        } else {
            uint32_t miss = mismatches(a, u, read_left + (pos - L), L, errors);
            if (miss == 0) {
                iv_push(path, cl.c[i].id);
                uint64_t nc = bo_str2num(u, k - 1);
                cands_free(&cl);
                return greedy_left(a, read, nc, pos - L, path, errors);
            } else if (miss < best) {
                ended = 0;
                best = miss;
                best_i = (uint32_t)i;
                next_len = cl.c[i].len;
                next_code = bo_str2num(u, k - 1);
            }
        }
    }
    if (best <= errors) {
        iv_push(path, cl.c[best_i].id);
        cands_free(&cl);
        if (ended) {
            iv_push(path, offset);
            return best;
        }
        return best + greedy_left(a, read, next_code, pos - (next_len - k + 1), path, errors - best);
    }
    cands_free(&cl);
    return best;
}

/* mapOnRightEndGreedy, alignerGreedy.cpp:221-265: later rightward steps. The remaining read
 * starts AT the overlap (so the k-1 overlap bases are compared again). */
static uint32_t greedy_right_more(bo_aligner* a, const char* read, uint32_t len, uint64_t code,
                                  uint32_t pos, ivec* path, uint32_t errors) {
    const uint32_t k = a->k;
    uint32_t rl = len - pos; /* |read.substr(pos)| */
    const char* read_left = read + pos;
    if (rl < k) return 0;
    cands_t cl;
    query(a, code, 0, &cl);
    uint32_t best = errors + 1, best_i = 9, next_len = 0;
    int ended = 0;
    uint64_t next_code = 0;
    for (int i = 0; i < cl.n; ++i) {
        const char* u = cl.c[i].seq;
        uint32_t ulen = cl.c[i].len;
        if (ulen - k + 1 >= rl) {
            uint32_t miss = mismatches(a, u, read_left, rl, errors);
            if (miss == 0) {
                iv_push(path, cl.c[i].id);
                cands_free(&cl);
                return 0;
            } else if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                ended = 1;
            }
        } else {
            /* whole unitig against read.substr(pos, |unitig|), cut at the end of the read */
            uint32_t n2 = ulen < rl ? ulen : rl;
            uint32_t miss = mismatches(a, u, read_left, n2, errors);
            if (miss == 0) {
                iv_push(path, cl.c[i].id);
                uint64_t nc = bo_str2num(u + ulen - (k - 1), k - 1);
                cands_free(&cl);
                return greedy_right_more(a, read, len, nc, pos + (ulen - k + 1), path, errors);
            } else if (miss < best) {
                ended = 0;
                best = miss;
                best_i = (uint32_t)i;
                next_len = ulen;
                next_code = bo_str2num(u + ulen - (k - 1), k - 1);
            }
        }
    }
    if (best <= errors) {
        iv_push(path, cl.c[best_i].id);
        cands_free(&cl);
        if (ended) return best;
        return best + greedy_right_more(a, read, len, next_code, pos + (next_len - k + 1), path, errors - best);
    }
    cands_free(&cl);
    return best;
}

/* checkEndGreedy, alignerGreedy.cpp:322-364: first rightward step; comparisons start after the
 * anchor overlap. Nothing right of the anchor -> 0 with nothing pushed. */
static uint32_t greedy_right_first(bo_aligner* a, const char* read, uint32_t len, uint64_t code,
                                   uint32_t pos, ivec* path, uint32_t errors) {
    const uint32_t k = a->k;
    uint32_t rl = len - (pos + k - 1);
    const char* read_left = read + pos + k - 1;
    if (rl == 0) return 0;
    cands_t cl;
    query(a, code, 0, &cl);
    uint32_t best = errors + 1, best_i = 9, next_len = 0;
    int ended = 0;
    uint64_t next_code = 0;
    for (int i = 0; i < cl.n; ++i) {
        const char* u = cl.c[i].seq;
        uint32_t ulen = cl.c[i].len;
        uint32_t L = ulen - (k - 1);
        if (L >= rl) {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, rl, errors);
            if (miss == 0) {
                iv_push(path, cl.c[i].id);
                cands_free(&cl);
                return 0;
            } else if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                ended = 1;
            }
        } else {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, L, errors);
            if (miss == 0) {
                iv_push(path, cl.c[i].id);
                uint64_t nc = bo_str2num(u + ulen - (k - 1), k - 1);
                cands_free(&cl);
                return greedy_right_more(a, read, len, nc, pos + L, path, errors);
            } else if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                next_code = bo_str2num(u + ulen - (k - 1), k - 1);
                next_len = ulen;
                ended = 0;
            }
        }
    }
    if (best <= errors) {
        iv_push(path, cl.c[best_i].id);
        cands_free(&cl);
        if (ended) return best;
        return best + greedy_right_more(a, read, len, next_code, pos + (next_len - k + 1), path, errors - best);
    }
    cands_free(&cl);
    return best;
}

/* alignReadGreedy, alignerGreedy.cpp:35-57. */
static void align_greedy(bo_aligner* a, const char* read, uint32_t len, int* overlap_found, uint32_t errors,
                         int* rc, ivec* out) {
    seeds_t seeds = {0, 0, 0};
    out->n = 0;
    scan_n_overlap(a, read, len, a->effort, &seeds);
    if (seeds.n == 0) {
        a->c[BO_C_NOOVERLAP]++;
        free(seeds.v);
        return;
    }
    *overlap_found = 1;
    ivec pb = {0, 0, 0}, pe = {0, 0, 0};
    for (uint32_t s = 0; s < seeds.n; ++s) {
        pb.n = 0;
        uint32_t eb = greedy_left(a, read, seeds.v[s].code, seeds.v[s].pos, &pb, errors);
        if (eb <= errors) {
            pe.n = 0;
            uint32_t ee = greedy_right_first(a, read, len, seeds.v[s].code, seeds.v[s].pos, &pe, errors - eb);
            if (ee + eb <= errors) {
                a->c[BO_C_ALIGNED]++;
                iv_reverse(&pb);
                iv_append(&pb, &pe);
                iv_assign(out, &pb);
                iv_free(&pb);
                iv_free(&pe);
                free(seeds.v);
                return;
            }
        }
    }
    iv_free(&pb);
    iv_free(&pe);
    free(seeds.v);
    if (!*rc) {
        *rc = 1;
        char* r = (char*)malloc((size_t)len + 1);
        bo_reverse_complement(read, len, r);
        r[len] = 0;
        align_greedy(a, r, len, overlap_found, errors, rc, out);
        free(r);
        return;
    }
    a->c[BO_C_NOTALIGNED]++;
}

/* ------------------------------------------------------------------ exhaustive extension
 * alignerExhaustive.cpp:61-259 (walks) and alignerPaths.cpp:127-175,289-460 (paths: a
 * unitig already present in `path` is skipped). `avoid` selects the path flavour. */

static uint32_t exh_right_more(bo_aligner* a, const char* read, uint32_t len, uint64_t code, uint32_t pos,
                               ivec* path, uint32_t errors, int avoid) {
    /* mapOnRightEndExhaustive :61-106 / mapOnRightEndExhaustivePath (alignerPaths.cpp:127-175) */
    const uint32_t k = a->k;
    uint32_t rl = len - (pos + k - 1);
    const char* read_left = read + pos + k - 1;
    if (rl == 0) {
        iv_push(path, 0);
        return 0;
    }
    cands_t cl;
    query(a, code, 0, &cl);
    uint32_t best = errors + 1, best_i = 9;
    int ended = 0;
    ivec keep = {0, 0, 0};
    for (int i = 0; i < cl.n; ++i) {
        if (avoid && iv_contains(path, cl.c[i].id)) continue;
        const char* u = cl.c[i].seq;
        uint32_t ulen = cl.c[i].len, L = ulen - (k - 1);
        if (rl <= L) {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, rl, errors);
            if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                ended = 1;
            }
        } else {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, L, errors);
            if (miss < best) {
                uint64_t nc = bo_str2num(u + ulen - (k - 1), k - 1);
                ivec possible = {0, 0, 0};
                miss += exh_right_more(a, read, len, nc, pos + L, &possible, errors - miss, avoid);
                if (miss < best) {
                    iv_assign(&keep, &possible);
                    best = miss;
                    best_i = (uint32_t)i;
                    ended = 0;
                }
                iv_free(&possible);
            }
        }
    }
    if (best <= errors) {
        iv_push(path, cl.c[best_i].id);
        if (ended) iv_push(path, (int32_t)(rl + k - 1));
        else iv_append(path, &keep);
    }
    iv_free(&keep);
    cands_free(&cl);
    return best;
}

static uint32_t exh_left_more(bo_aligner* a, const char* read, uint64_t code, uint32_t pos, ivec* path,
                              uint32_t errors) {
    /* mapOnLeftEndExhaustive :109-155 (also what mapOnLeftEndExhaustivePaths recurses into,
     * alignerPaths.cpp:440). Ends exactly on a unitig start: nothing pushed. */
    const uint32_t k = a->k;
    if (pos == 0) return 0;
    const char* read_left = read;
    cands_t cl;
    query(a, code, 1, &cl);
    uint32_t best = errors + 1, best_i = 9;
    int32_t offset = -2;
    int ended = 0;
    ivec keep = {0, 0, 0};
    for (int i = 0; i < cl.n; ++i) {
        const char* u = cl.c[i].seq;
        uint32_t ulen = cl.c[i].len, L = ulen - (k - 1);
        if (pos + k - 1 <= ulen) {
            uint32_t miss = mismatches(a, u + (L - pos), read_left, pos, errors);
            if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                offset = (int32_t)(L - pos);
                ended = 1;
            }
        } else {
            uint32_t miss = mismatches(a, u, read_left + (pos - L), L, errors);
            if (miss < best) {
                uint64_t nc = bo_str2num(u, k - 1);
                ivec possible = {0, 0, 0};
                miss += exh_left_more(a, read, nc, pos - L, &possible, errors - miss);
                if (miss < best) {
                    iv_assign(&keep, &possible);
                    best = miss;
                    best_i = (uint32_t)i;
                    offset = -1;
                    ended = 0;
                }
                iv_free(&possible);
            }
        }
    }
    if (best <= errors) {
        if (ended) iv_push(path, offset);
        else iv_append(path, &keep);
        iv_push(path, cl.c[best_i].id);
    }
    iv_free(&keep);
    cands_free(&cl);
    return best;
}

static uint32_t exh_left_first(bo_aligner* a, const char* read, uint64_t code, uint32_t pos, ivec* path,
                               uint32_t errors, int avoid) {
    /* checkBeginExhaustive :158-203 / checkBeginExhaustivePath (alignerPaths.cpp:351-411) */
    const uint32_t k = a->k;
    if (pos == 0) {
        iv_push(path, 0);
        return 0;
    }
    const char* read_left = read;
    cands_t cl;
    query(a, code, 1, &cl);
    uint32_t best = errors + 1, best_i = 0;
    int32_t offset = -2;
    int ended = 0;
    ivec keep = {0, 0, 0};
    for (int i = 0; i < cl.n; ++i) {
        if (avoid && iv_contains(path, cl.c[i].id)) continue;
        const char* u = cl.c[i].seq;
        uint32_t ulen = cl.c[i].len, L = ulen - (k - 1);
        if (L >= pos) {
            uint32_t miss = mismatches(a, u + (L - pos), read_left, pos, errors);
            if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                ended = 1;
                offset = (int32_t)(L - pos);
            }
        } else {
            uint32_t miss = mismatches(a, u, read_left + (pos - L), L, errors);
            if (miss < best) {
                uint64_t nc = bo_str2num(u, k - 1);
                ivec possible = {0, 0, 0};
                /* the path flavour calls mapOnLeftEndExhaustivePaths, whose body equals
                 * mapOnLeftEndExhaustive (alignerPaths.cpp:414-460) */
                miss += exh_left_more(a, read, nc, pos - L, &possible, errors - miss);
                if (miss < best) {
                    best = miss;
                    best_i = (uint32_t)i;
                    iv_assign(&keep, &possible);
                    ended = 0;
                }
                iv_free(&possible);
            }
        }
    }
    if (best <= errors) {
        if (ended) {
            /* path flavour re-checks membership before pushing (alignerPaths.cpp:400-404) */
            if (!avoid || !iv_contains(path, cl.c[best_i].id)) {
                iv_push(path, offset);
                iv_push(path, cl.c[best_i].id);
            }
        } else {
            iv_append(path, &keep);
            iv_push(path, cl.c[best_i].id);
        }
    }
    iv_free(&keep);
    cands_free(&cl);
    return best;
}

static uint32_t exh_right_first(bo_aligner* a, const char* read, uint32_t len, uint64_t code, uint32_t pos,
                                ivec* path, uint32_t errors, int avoid) {
    /* checkEndExhaustive :206-259 / checkEndExhaustivePath (alignerPaths.cpp:289-348);
     * partial mode (-i) is out of scope and off. */
    const uint32_t k = a->k;
    uint32_t rl = len - (pos + k - 1);
    const char* read_left = read + pos + k - 1;
    if (rl == 0) {
        iv_push(path, 0);
        return 0;
    }
    cands_t cl;
    query(a, code, 0, &cl);
    uint32_t best = errors + 1, best_i = 9;
    int ended = 0;
    int32_t offset = -2;
    ivec keep = {0, 0, 0};
    for (int i = 0; i < cl.n; ++i) {
        if (avoid && iv_contains(path, cl.c[i].id)) continue;
        const char* u = cl.c[i].seq;
        uint32_t ulen = cl.c[i].len, L = ulen - (k - 1);
        if (L >= rl) {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, rl, errors);
            if (miss < best) {
                best = miss;
                best_i = (uint32_t)i;
                ended = 1;
                offset = (int32_t)(rl + k - 1);
            }
        } else {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, L, errors);
            if (miss < best) {
                uint64_t nc = bo_str2num(u + ulen - (k - 1), k - 1);
                ivec possible = {0, 0, 0};
                miss += exh_right_more(a, read, len, nc, pos + L, &possible, errors - miss, avoid);
                if (miss < best) {
                    iv_assign(&keep, &possible);
                    best = miss;
                    best_i = (uint32_t)i;
                    ended = 0;
                }
                iv_free(&possible);
            }
        }
    }
    if (best <= errors) {
        if (ended) {
            if (!avoid || !iv_contains(path, cl.c[best_i].id)) {
                iv_push(path, cl.c[best_i].id);
                iv_push(path, offset);
            }
        } else {
            iv_push(path, cl.c[best_i].id);
            iv_append(path, &keep);
        }
    }
    iv_free(&keep);
    cands_free(&cl);
    return best;
}

/* alignReadExhaustive (alignerExhaustive.cpp:35-58) and alignReadExhaustivePath
 * (alignerPaths.cpp:66-88): same driver, the latter on the path-flavoured helpers. */
static void align_exhaustive(bo_aligner* a, const char* read, uint32_t len, int* overlap_found, uint32_t errors,
                             ivec* out, int avoid) {
    seeds_t seeds = {0, 0, 0};
    out->n = 0;
    scan_all(a, read, len, &seeds);
    a->c[BO_C_OVERLAPS] += seeds.n; /* the list is never empty: window 0 always exists */
    *overlap_found = 1;
    ivec pb = {0, 0, 0}, pe = {0, 0, 0};
    for (uint32_t s = 0; s < seeds.n; ++s) {
        pb.n = 0;
        uint32_t eb = exh_left_first(a, read, seeds.v[s].code, seeds.v[s].pos, &pb, errors, avoid);
        if (eb <= errors) {
            pe.n = 0;
            uint32_t ee = exh_right_first(a, read, len, seeds.v[s].code, seeds.v[s].pos, &pe, errors - eb, avoid);
            if (ee + eb <= errors) {
                a->c[BO_C_ALIGNED]++;
                a->c[BO_C_READS]++;
                iv_append(&pb, &pe);
                iv_assign(out, &pb);
                iv_free(&pb);
                iv_free(&pe);
                free(seeds.v);
                return;
            }
        }
    }
    a->c[BO_C_NOTALIGNED]++;
    a->c[BO_C_READS]++;
    iv_free(&pb);
    iv_free(&pe);
    free(seeds.v);
}

/* ------------------------------------------------------------------ greedy "paths" aligner
 * alignerPaths.cpp:35-63,91-124,178-286. */

static uint32_t check_pair_paths(bo_aligner* a, const seed_t* o1, const seed_t* o2, const char* read, int32_t* number,
                                 uint32_t allowed, const ivec* path) {
    /* checkPairPaths, alignerPaths.cpp:237-286 */
    const uint32_t k = a->k;
    cands_t c1, c2;
    if (o2->pos - o1->pos < k) {
        query(a, o1->code, 0, &c1);
        query(a, o2->code, 1, &c2);
        for (int i = 0; i < c1.n; ++i) {
            for (int j = 0; j < c2.n; ++j) {
                if (c2.c[j].id == c1.c[i].id) {
                    if (!iv_contains(path, c1.c[i].id)) *number = c1.c[i].id;
                    cands_free(&c1);
                    cands_free(&c2);
                    return 0;
                }
            }
        }
        cands_free(&c1);
        cands_free(&c2);
        return allowed + 1;
    }
    uint32_t sub_len = o2->pos - o1->pos - (k - 1);
    const char* sub = read + o1->pos + k - 1;
    query(a, o1->code, 0, &c1);
    query(a, o2->code, 1, &c2);
    uint32_t best = allowed + 1, best_i = 0;
    for (int i = 0; i < c1.n; ++i) {
        for (int j = 0; j < c2.n; ++j) {
            if (c2.c[j].id == c1.c[i].id && !iv_contains(path, c1.c[i].id)) {
                /* getUnitig(id): the oriented unitig again (aligner.cpp:293-302) */
                if ((uint64_t)c1.c[i].len - 2 * (uint64_t)(k - 1) == sub_len) {
                    uint32_t miss = mismatches(a, c1.c[i].seq + (k - 1), sub, sub_len, allowed);
                    if (miss < best) {
                        best = miss;
                        best_i = (uint32_t)i;
                    }
                }
            }
        }
    }
    if (best <= allowed) *number = c1.c[best_i].id;
    cands_free(&c1);
    cands_free(&c2);
    return best;
}

typedef struct {
    uint32_t first, second;
} upair;

static upair map_on_right_path(bo_aligner* a, const char* read, uint32_t len, ivec* path, uint64_t code, uint32_t pos,
                               const seeds_t* list, int* ended, uint32_t start, uint32_t errors) {
    /* mapOnRightPath, alignerPaths.cpp:178-234 */
    const uint32_t k = a->k;
    uint32_t rl = len - (pos + k - 1);
    const char* read_left = read + pos + k - 1;
    upair r;
    if (rl == 0) { /* the reference prints "should not appears" and exit(0)s here */
        r.first = start;
        r.second = 0;
        return r;
    }
    cands_t cl;
    query(a, code, 0, &cl);
    uint32_t best = errors + 1, best_i = 9, next = start, next_len = 0;
    uint64_t next_code = 0;
    for (int i = 0; i < cl.n; ++i) {
        if (iv_contains(path, cl.c[i].id)) continue;
        const char* u = cl.c[i].seq;
        uint32_t ulen = cl.c[i].len, L = ulen - (k - 1);
        if (rl <= L) {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, rl, errors);
            if (miss < best) {
                *ended = 1;
                best = miss;
                best_i = (uint32_t)i;
            }
        } else {
            uint32_t miss = mismatches(a, u + (k - 1), read_left, L, errors);
            if (miss < best) {
                uint64_t oc = bo_str2num(u + ulen - (k - 1), k - 1);
                *ended = 0;
                best = miss;
                best_i = (uint32_t)i;
                next_code = oc;
                next_len = ulen;
                next = start;
                for (uint32_t j = start + 1; j < list->n; ++j)
                    if (oc == list->v[j].code && list->v[j].pos == pos + ulen - k + 1) next = j;
            }
        }
    }
    if (*ended) {
        /* NB the reference indexes rangeUnitigs[9] here when `ended` was already true on entry
         * and no candidate improved; callers always pass ended=false on first entry. */
        if (best_i < (uint32_t)cl.n) iv_push(path, cl.c[best_i].id);
        cands_free(&cl);
        r.first = start;
        r.second = best;
        return r;
    }
    if (best <= errors) {
        iv_push(path, cl.c[best_i].id);
        cands_free(&cl);
        if (next > start) {
            r.first = next;
            r.second = best;
            return r;
        }
        upair s = map_on_right_path(a, read, len, path, next_code, pos + (next_len - k + 1), list, ended, start,
                                    errors - best);
        r.first = s.first;
        r.second = s.second + best;
        return r;
    }
    cands_free(&cl);
    r.first = start;
    r.second = errors + 1;
    return r;
}

static uint32_t cover_greedy_path(bo_aligner* a, const char* read, uint32_t len, const seeds_t* list, uint32_t start,
                                  uint32_t end, ivec* path, uint32_t errors, int* ended) {
    /* coverGreedyPath, alignerPaths.cpp:91-124 */
    if (start == end) return 0;
    int indice = 0;
    uint32_t best = errors + 1;
    int32_t keep = 0;
    for (uint32_t i = 1; i + start <= end && i <= a->effort; ++i) {
        int32_t number = 0;
        uint32_t miss = check_pair_paths(a, &list->v[start], &list->v[i + start], read, &number, errors, path);
        if (miss < best) {
            indice = (int)i;
            keep = number;
            best = miss;
        }
    }
    if (best <= errors) {
        iv_push(path, keep);
        return best + cover_greedy_path(a, read, len, list, (uint32_t)indice + start, end, path, errors - best, ended);
    }
    upair p = map_on_right_path(a, read, len, path, list->v[start].code, list->v[start].pos, list, ended, start, errors);
    if (p.second <= errors) {
        if (*ended) return p.second;
        if (p.first > start) {
            uint32_t rec = cover_greedy_path(a, read, len, list, p.first, end, path, errors - p.second, ended);
            if (rec + p.second <= errors) return rec + p.second;
        }
    }
    return errors + 1;
}

static void align_greedy_path(bo_aligner* a, const char* read, uint32_t len, int* overlap_found, uint32_t errors,
                              ivec* out) {
    /* alignReadGreedyPath, alignerPaths.cpp:35-63 */
    seeds_t list = {0, 0, 0};
    out->n = 0;
    scan_all(a, read, len, &list);
    a->c[BO_C_OVERLAPS] += list.n;
    *overlap_found = 1;
    uint32_t try_n = a->effort;
    uint32_t last_start = try_n - 1 < list.n - 1 ? try_n - 1 : list.n - 1;
    for (uint32_t start = 0; start <= last_start; ++start) {
        ivec path = {0, 0, 0};
        uint32_t eb = exh_left_first(a, read, list.v[start].code, list.v[start].pos, &path, errors, 1);
        if (eb <= errors) {
            int lo = (int)start > (int)list.n - (int)try_n ? (int)start : (int)list.n - (int)try_n;
            for (int end = (int)list.n - 1; end >= lo; --end) {
                uint32_t ee = exh_right_first(a, read, len, list.v[end].code, list.v[end].pos, &path, errors - eb, 1);
                if (ee + eb <= errors) {
                    int ended = 0;
                    uint32_t ec = cover_greedy_path(a, read, len, &list, start, (uint32_t)end, &path,
                                                    errors - ee - eb, &ended);
                    if (ec <= errors - ee - eb) {
                        a->c[BO_C_ALIGNED]++;
                        a->c[BO_C_READS]++;
                        iv_assign(out, &path);
                        iv_free(&path);
                        free(list.v);
                        return;
                    }
                }
            }
        }
        iv_free(&path);
    }
    a->c[BO_C_NOTALIGNED]++;
    a->c[BO_C_READS]++;
    free(list.v);
}

/* ------------------------------------------------------------------ per-read entry */

int bo_align(bo_aligner* a, int mode, const char* read, uint32_t len, int* overlap_found, int* rc, int32_t* path,
             int cap) {
    ivec out = {0, 0, 0};
    *overlap_found = 0;
    *rc = 0;
    if (len + 1 < a->k) {
        /* Shorter than one (k-1)-mer window (only a FASTQ input can deliver this, aligner.cpp:54).
         * The reference either finds nothing or dies in substr (std::out_of_range); the defined
         * outcome chosen here and in the product: the read has no overlap / is not aligned. */
        a->c[BO_C_READS]++;
        if (mode == BO_GREEDY) a->c[BO_C_NOOVERLAP]++;
        else a->c[BO_C_NOTALIGNED]++;
        return 0;
    }
    switch (mode) {
        case BO_GREEDY:
            a->c[BO_C_READS]++; /* alignerGreedy.cpp:382 */
            align_greedy(a, read, len, overlap_found, a->errors, rc, &out);
            break;
        case BO_EXHAUSTIVE: align_exhaustive(a, read, len, overlap_found, a->errors, &out, 0); break;
        case BO_EXHAUSTIVE_PATH: align_exhaustive(a, read, len, overlap_found, a->errors, &out, 1); break;
        case BO_GREEDY_PATH: align_greedy_path(a, read, len, overlap_found, a->errors, &out); break;
        default: return -1;
    }
    int n = out.n;
    if (n > cap) {
        iv_free(&out);
        return -1;
    }
    memcpy(path, out.v, sizeof(int32_t) * (size_t)n);
    iv_free(&out);
    return n;
}

/* ------------------------------------------------------------------ read parsing
 * getReads, aligner.cpp:46-117. */

typedef struct {
    sbuf hdr, seq;
    uint64_t *hoff, *soff;
    uint64_t n, cap;
} readset_t;

static void rs_push(readset_t* rs, const sbuf* h, const sbuf* r) {
    if (rs->n + 1 >= rs->cap) {
        rs->cap = rs->cap ? 2 * rs->cap : 1024;
        rs->hoff = (uint64_t*)realloc(rs->hoff, sizeof(uint64_t) * (rs->cap + 1));
        rs->soff = (uint64_t*)realloc(rs->soff, sizeof(uint64_t) * (rs->cap + 1));
    }
    if (rs->n == 0) rs->hoff[0] = rs->soff[0] = 0;
    sb_put(&rs->hdr, h->s, h->n);
    sb_put(&rs->seq, r->s, r->n);
    rs->n++;
    rs->hoff[rs->n] = rs->hdr.n;
    rs->soff[rs->n] = rs->seq.n;
}

static int acgtn_only(const sbuf* r) {
    for (size_t j = 0; j < r->n; ++j) {
        char c = r->s[j];
        if (c != 'A' && c != 'C' && c != 'T' && c != 'G' && c != 'N') return 0;
    }
    return 1;
}

/* one getReads(reads, n) call; strings are locals of the call, hence fresh every time */
static void get_reads(istream_t* s, uint32_t k, int fastq, uint32_t n, readset_t* rs) {
    sbuf read = {0, 0, 0}, header = {0, 0, 0}, inter = {0, 0, 0};
    if (fastq) {
        for (uint32_t i = 0; i < n; ++i) {
            is_getline(s, &header);
            is_getline(s, &read);
            if (read.n > 2 && acgtn_only(&read)) rs_push(rs, &header, &read);
            is_getline(s, &header); /* '+' line and quality line both land in `header` */
            is_getline(s, &header);
            if (s->eofbit) break;
        }
    } else {
        for (uint32_t i = 0; i < n; ++i) {
            is_getline(s, &header);
            is_getline(s, &read);
            for (;;) {
                int c = is_peek(s);
                if ((char)c == '>') {
                    if (read.n > 2 && acgtn_only(&read) && read.n > k) rs_push(rs, &header, &read);
                    read.n = 0;
                    break;
                } else if (!s->eofbit) {
                    is_getline(s, &inter);
                    sb_put(&read, inter.s, inter.n);
                } else {
                    if (read.n > 2 && acgtn_only(&read) && read.n > k) rs_push(rs, &header, &read);
                    goto done;
                }
            }
        }
    }
done:
    free(read.s);
    free(header.s);
    free(inter.s);
}

static int parse_file(const char* path, uint32_t k, int fastq, uint32_t chunk, readset_t* rs) {
    size_t n = 0;
    char* data = slurp(path, &n);
    if (!data) return -1;
    istream_t s = {data, n, 0, 0, 0};
    while (!s.eofbit) get_reads(&s, k, fastq, chunk, rs); /* alignerGreedy.cpp:372-377 */
    free(data);
    return 0;
}

int64_t bo_parse_reads(const char* reads_path, uint32_t k, int fastq, uint32_t chunk, char** headers,
                       uint64_t** header_off, char** reads, uint64_t** read_off) {
    readset_t rs;
    memset(&rs, 0, sizeof rs);
    if (parse_file(reads_path, k, fastq, chunk, &rs) != 0) return -1;
    if (rs.n == 0) {
        rs.hoff = (uint64_t*)calloc(1, sizeof(uint64_t));
        rs.soff = (uint64_t*)calloc(1, sizeof(uint64_t));
        sb_reserve(&rs.hdr, 1);
        sb_reserve(&rs.seq, 1);
    }
    *headers = rs.hdr.s;
    *header_off = rs.hoff;
    *reads = rs.seq.s;
    *read_off = rs.soff;
    return (int64_t)rs.n;
}

void bo_free_buf(void* p) { free(p); }

/* ------------------------------------------------------------------ whole run */

/* printPath, aligner.cpp:600-609 */
static void print_path(sbuf* out, const int32_t* path, int n) {
    char tmp[16];
    for (int i = 0; i < n; ++i) {
        int m = snprintf(tmp, sizeof tmp, "%d.", path[i]);
        sb_put(out, tmp, (size_t)m);
    }
    sb_putc(out, '\n');
}

/* ---- correction mode (-c): recoverPath / getUnitig (aligner.cpp:268-302), compactionEnd (utils.cpp:171-179) ---- */

static void oriented_unitig(const bo_aligner* a, int32_t id, sbuf* out) { /* getUnitig, aligner.cpp:293-302 */
    out->n = 0;
    uint32_t u = (uint32_t)(id > 0 ? id : -id);
    if (u >= a->n_unitigs) return; /* the reference indexes its vector out of range here */
    if (id > 0) {
        sb_put(out, a->unitigs[u], a->ulen[u]);
    } else {
        char* tmp = (char*)malloc(a->ulen[u] + 1);
        bo_reverse_complement(a->unitigs[u], a->ulen[u], tmp);
        sb_put(out, tmp, a->ulen[u]);
        free(tmp);
    }
}

/* recoverPath(numbers, size) followed by the reverse complement alignPartGreedy applies when rc is set
 * (alignerGreedy.cpp:394-398). Returns the corrected read's length, -1 where the reference prints
 * "bug compaction" and exits, -2 where it would index numbers[1] of a one-element path. */
int64_t bo_recover_path(const bo_aligner* a, const int32_t* numbers, int n, uint32_t size, int rc, char* out) {
    if (n < 2) return -2;
    const uint32_t k1 = a->k - 1;
    sbuf path = {0, 0, 0}, un = {0, 0, 0};
    oriented_unitig(a, numbers[1], &path);
    for (int i = 2; i < n; ++i) {
        oriented_unitig(a, numbers[i], &un);
        /* compactionEnd(path, unitig, k-1) */
        if (path.n == 0 || un.n == 0 || path.n < k1 || un.n < k1) {
            free(path.s);
            free(un.s);
            return -1;
        }
        const char* end1 = path.s + path.n - k1;
        if (memcmp(end1, un.s, k1) == 0) {
            sb_put(&path, un.s + k1, un.n - k1);
        } else {
            char* r2 = (char*)malloc(un.n + 1);
            bo_reverse_complement(un.s, (uint32_t)un.n, r2);
            if (memcmp(end1, r2, k1) == 0) {
                sb_put(&path, r2 + k1, un.n - k1);
                free(r2);
            } else {
                free(r2);
                free(path.s);
                free(un.s);
                return -1;
            }
        }
    }
    /* path.substr(offset, size) */
    const int64_t offset = numbers[0];
    int64_t len = 0;
    if (offset >= 0 && (uint64_t)offset <= path.n) {
        len = (int64_t)path.n - offset;
        if (len > (int64_t)size) len = size;
        if (rc) bo_reverse_complement(path.s + offset, (uint32_t)len, out);
        else memcpy(out, path.s + offset, (size_t)len);
    } else {
        len = -1; /* std::out_of_range in the reference */
    }
    free(path.s);
    free(un.s);
    return len;
}

static int run_impl(bo_aligner* a, int mode, const char* reads_csv, int fastq, const char* paths_file,
                    const char* not_aligned_file, int correction);

int bo_run(bo_aligner* a, int mode, const char* reads_csv, int fastq, const char* paths_file,
           const char* not_aligned_file) {
    return run_impl(a, mode, reads_csv, fastq, paths_file, not_aligned_file, 0);
}

int bo_run_correct(bo_aligner* a, const char* reads_csv, int fastq, const char* paths_file,
                   const char* not_aligned_file) {
    return run_impl(a, BO_GREEDY, reads_csv, fastq, paths_file, not_aligned_file, 1);
}

static int run_impl(bo_aligner* a, int mode, const char* reads_csv, int fastq, const char* paths_file,
                    const char* not_aligned_file, int correction) {
    FILE* fp = fopen(paths_file, "wb");
    FILE* fn = fopen(not_aligned_file, "wb");
    if (!fp || !fn) return -1;
    int rcode = 0;
    const char* p = reads_csv;
    int32_t* path = (int32_t*)malloc(sizeof(int32_t) * (1 << 20));
    for (;;) {
        const char* comma = strchr(p, ',');
        size_t fl = comma ? (size_t)(comma - p) : strlen(p);
        char* file = (char*)malloc(fl + 1);
        memcpy(file, p, fl);
        file[fl] = 0;
        readset_t rs;
        memset(&rs, 0, sizeof rs);
        if (parse_file(file, a->k, fastq, 10000, &rs) != 0) rcode = -1;
        for (uint64_t i = 0; i < rs.n; ++i) {
            const char* h = rs.hdr.s + rs.hoff[i];
            size_t hn = rs.hoff[i + 1] - rs.hoff[i];
            const char* r = rs.seq.s + rs.soff[i];
            uint32_t rn = (uint32_t)(rs.soff[i + 1] - rs.soff[i]);
            int of = 0, rc = 0;
            int n = bo_align(a, mode, r, rn, &of, &rc, path, 1 << 20);
            sbuf rec = {0, 0, 0};
            sb_put(&rec, h, hn);
            sb_putc(&rec, '\n');
            if (n > 0 && correction) { /* alignerGreedy.cpp:394-404 */
                char* corr = (char*)malloc((size_t)rn + 1);
                int64_t cl = bo_recover_path(a, path, n, rn, rc, corr);
                if (cl < 0) { /* the reference exits / throws here */
                    free(corr);
                    free(rec.s);
                    rcode = -2;
                    break;
                }
                sb_put(&rec, corr, (size_t)cl);
                sb_putc(&rec, '\n');
                fwrite(rec.s, 1, rec.n, fp);
                free(corr);
            } else if (n > 0) { /* alignerGreedy.cpp:405-411 */
                print_path(&rec, path, n);
                fwrite(rec.s, 1, rec.n, fp);
            } else { /* alignerGreedy.cpp:420-426: the original read, never the RC one */
                sb_put(&rec, r, rn);
                sb_putc(&rec, '\n');
                fwrite(rec.s, 1, rec.n, fn);
            }
            free(rec.s);
        }
        free(rs.hdr.s);
        free(rs.seq.s);
        free(rs.hoff);
        free(rs.soff);
        free(file);
        if (!comma) break;
        p = comma + 1;
    }
    free(path);
    fclose(fp);
    fclose(fn);
    return rcode;
}
