/*
 * oracle/bm25_oracle.c -- CPU restatement of the BM25 scorer behind
 * TextSearch()/HybridSearch().
 *
 * TEST INFRASTRUCTURE ONLY (see vs_oracle.c header).
 *
 * The reference scores through TANTIVY::ffi_bm25_search
 * (src/Storages/MergeTree/TantivyIndexStore.cpp:900-954), implemented in the
 * un-vendored Rust crate tantivy_search 0.1.0 over tantivy 0.21.1
 * (rust/supercrate/Cargo.lock:2077-2079, :2202-2203).  The algorithm restated
 * here is tantivy 0.21's published BM25:
 *     k1 = 1.2, b = 0.75
 *     idf    = ln(1 + (N - n + 0.5) / (n + 0.5))                     (fp32)
 *     weight = idf * (1 + k1)
 *     norm[c]= k1 * (1 - b + b * fieldnorm(c) / avgdl),  avgdl = total_tokens / total_docs (fp32)
 *     score  = weight * tf / (tf + norm[fieldnorm_id(dl)])
 * with tantivy's 1-byte field-norm code (exact for dl < 40) and the "default"
 * tokenizer (split on non-alphanumeric, drop tokens of 40 bytes or more, lowercase).
 * Multi-term queries are a Boolean OR (operator_or) / AND of term scorers whose
 * scores add; top-k ties go to the smaller doc id (TopDocs).
 * Table-wide statistics override the per-part ones exactly like
 * ReadWithHybridSearch::getStatisticForTextSearch
 * (src/VectorIndex/Processors/ReadWithHybridSearch.cpp:89-209) +
 * BM25InfoInDataParts (src/VectorIndex/Common/BM25InfoInDataParts.cpp:40-94).
 * Pinned by goldens 00040_mqvs_hybrid_search.reference (2.1646233, 1.9431154,
 * 2.7369592, 0.9453843, 0.89381784, 0.8700882, 2.2973092) in
 * tests/test_oracle_golden.py.
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BM25_K1 1.2f
#define BM25_B 0.75f

static uint32_t g_fieldnorm_table[256];
static int g_fieldnorm_init = 0;

/* tantivy fieldnorm/code.rs FIELD_NORMS_TABLE: 0..39 exact, then 8 codes per
 * doubling step (40,42,..,54, 56,60,..,84, 88,96,..). */
static void fieldnorm_init(void) {
    if (g_fieldnorm_init) return;
    for (int i = 0; i < 40; i++) g_fieldnorm_table[i] = (uint32_t)i;
    uint64_t v = 40;
    for (int i = 40; i < 256; i++) {
        g_fieldnorm_table[i] = v > 0xffffffffull ? 0xffffffffu : (uint32_t)v;
        int shift = (i - 40) / 8 + 1;
        v += shift >= 40 ? (1ull << 40) : (1ull << shift);
    }
    g_fieldnorm_init = 1;
}

uint32_t orc_bm25_id_to_fieldnorm(int id) {
    fieldnorm_init();
    return g_fieldnorm_table[id & 255];
}

int orc_bm25_fieldnorm_to_id(uint32_t n) {
    fieldnorm_init();
    int lo = 0, hi = 255;
    while (lo < hi) { /* largest id with table[id] <= n */
        int mid = (lo + hi + 1) / 2;
        if (g_fieldnorm_table[mid] <= n)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

typedef struct {
    uint32_t doc, tf;
} posting;

typedef struct {
    char *term;
    uint32_t field;
    posting *p;
    uint32_t n, cap;
} term_entry;

typedef struct orc_bm25_index {
    uint32_t n_fields;
    term_entry *terms;
    uint32_t n_terms, cap_terms;
    int64_t *slots; /* hash -> term idx */
    uint32_t n_slots;
    uint32_t n_docs, cap_docs;
    uint32_t *doc_len; /* [n_fields][cap_docs] token counts */
    uint64_t *row_id;
    uint64_t *total_tokens; /* per field */
} orc_bm25_index;

static uint64_t hash_term(uint32_t field, const char *s, size_t len) {
    uint64_t h = 1469598103934665603ull ^ field;
    for (size_t i = 0; i < len; i++) {
        h ^= (uint8_t)s[i];
        h *= 1099511628211ull;
    }
    return h;
}

orc_bm25_index *orc_bm25_create(uint32_t n_fields) {
    orc_bm25_index *ix = (orc_bm25_index *)calloc(1, sizeof(*ix));
    ix->n_fields = n_fields;
    ix->n_slots = 1u << 16;
    ix->slots = (int64_t *)malloc(sizeof(int64_t) * ix->n_slots);
    for (uint32_t i = 0; i < ix->n_slots; i++) ix->slots[i] = -1;
    ix->total_tokens = (uint64_t *)calloc(n_fields, sizeof(uint64_t));
    return ix;
}

void orc_bm25_free(orc_bm25_index *ix) {
    if (!ix) return;
    for (uint32_t i = 0; i < ix->n_terms; i++) {
        free(ix->terms[i].term);
        free(ix->terms[i].p);
    }
    free(ix->terms);
    free(ix->slots);
    free(ix->doc_len);
    free(ix->row_id);
    free(ix->total_tokens);
    free(ix);
}

static void rehash(orc_bm25_index *ix) {
    uint32_t ns = ix->n_slots * 2;
    int64_t *s = (int64_t *)malloc(sizeof(int64_t) * ns);
    for (uint32_t i = 0; i < ns; i++) s[i] = -1;
    for (uint32_t t = 0; t < ix->n_terms; t++) {
        uint64_t h = hash_term(ix->terms[t].field, ix->terms[t].term, strlen(ix->terms[t].term));
        uint32_t j = (uint32_t)(h & (ns - 1));
        while (s[j] >= 0) j = (j + 1) & (ns - 1);
        s[j] = t;
    }
    free(ix->slots);
    ix->slots = s;
    ix->n_slots = ns;
}

static int64_t find_term(const orc_bm25_index *ix, uint32_t field, const char *s, size_t len) {
    uint64_t h = hash_term(field, s, len);
    uint32_t j = (uint32_t)(h & (ix->n_slots - 1));
    while (ix->slots[j] >= 0) {
        const term_entry *e = &ix->terms[ix->slots[j]];
        if (e->field == field && strlen(e->term) == len && memcmp(e->term, s, len) == 0) return ix->slots[j];
        j = (j + 1) & (ix->n_slots - 1);
    }
    return -1;
}

static int64_t intern_term(orc_bm25_index *ix, uint32_t field, const char *s, size_t len) {
    int64_t t = find_term(ix, field, s, len);
    if (t >= 0) return t;
    if ((ix->n_terms + 1) * 2 > ix->n_slots) rehash(ix);
    if (ix->n_terms == ix->cap_terms) {
        ix->cap_terms = ix->cap_terms ? ix->cap_terms * 2 : 1024;
        ix->terms = (term_entry *)realloc(ix->terms, sizeof(term_entry) * ix->cap_terms);
    }
    term_entry *e = &ix->terms[ix->n_terms];
    memset(e, 0, sizeof(*e));
    e->term = (char *)malloc(len + 1);
    memcpy(e->term, s, len);
    e->term[len] = 0;
    e->field = field;
    uint64_t h = hash_term(field, s, len);
    uint32_t j = (uint32_t)(h & (ix->n_slots - 1));
    while (ix->slots[j] >= 0) j = (j + 1) & (ix->n_slots - 1);
    ix->slots[j] = ix->n_terms;
    return ix->n_terms++;
}

/* tantivy 0.21 "default" analyzer: SimpleTokenizer (maximal runs of chars with char::is_alphanumeric) ->
 * RemoveLongFilter::limit(40) (keeps tokens of FEWER than 40 bytes: `token.text.len() < limit`) -> LowerCaser
 * (Unicode to_lowercase).  Text is UTF-8.  The Unicode classes are restated for the blocks that matter in practice:
 * Latin-1 / General Punctuation / symbols, arrows, CJK and full-width punctuation are separators; Latin-1, Latin
 * Extended-A, Greek, Cyrillic and full-width capitals are lowercased; every other non-ASCII code point is a letter
 * and is kept as is.  (Parity beyond these blocks is unpinned)  Calls cb(token, len). */
typedef void (*tok_cb)(void *ctx, const char *tok, size_t len);
static size_t utf8_decode(const unsigned char *s, size_t n, uint32_t *cp) {
    if (s[0] < 0x80) { *cp = s[0]; return 1; }
    if ((s[0] & 0xE0) == 0xC0 && n >= 2 && (s[1] & 0xC0) == 0x80) { *cp = ((s[0] & 0x1Fu) << 6) | (s[1] & 0x3Fu); return 2; }
    if ((s[0] & 0xF0) == 0xE0 && n >= 3 && (s[1] & 0xC0) == 0x80 && (s[2] & 0xC0) == 0x80) {
        *cp = ((s[0] & 0x0Fu) << 12) | ((s[1] & 0x3Fu) << 6) | (s[2] & 0x3Fu); return 3;
    }
    if ((s[0] & 0xF8) == 0xF0 && n >= 4 && (s[1] & 0xC0) == 0x80 && (s[2] & 0xC0) == 0x80 && (s[3] & 0xC0) == 0x80) {
        *cp = ((s[0] & 0x07u) << 18) | ((s[1] & 0x3Fu) << 12) | ((s[2] & 0x3Fu) << 6) | (s[3] & 0x3Fu); return 4;
    }
    *cp = 0xFFFD; /* invalid byte: a letter-like replacement, one byte consumed */
    return 1;
}
static size_t utf8_encode(uint32_t cp, char *out) {
    if (cp < 0x80) { out[0] = (char)cp; return 1; }
    if (cp < 0x800) { out[0] = (char)(0xC0 | (cp >> 6)); out[1] = (char)(0x80 | (cp & 0x3F)); return 2; }
    if (cp < 0x10000) { out[0] = (char)(0xE0 | (cp >> 12)); out[1] = (char)(0x80 | ((cp >> 6) & 0x3F)); out[2] = (char)(0x80 | (cp & 0x3F)); return 3; }
    out[0] = (char)(0xF0 | (cp >> 18)); out[1] = (char)(0x80 | ((cp >> 12) & 0x3F)); out[2] = (char)(0x80 | ((cp >> 6) & 0x3F));
    out[3] = (char)(0x80 | (cp & 0x3F)); return 4;
}
static int cp_is_alnum(uint32_t c) {
    if (c < 0x80) return isalnum((int)c) != 0;
    if (c <= 0xBF) return c == 0xAA || c == 0xB2 || c == 0xB3 || c == 0xB5 || c == 0xB9 || c == 0xBA || c == 0xBC || c == 0xBD || c == 0xBE;
    if (c == 0xD7 || c == 0xF7) return 0;
    if (c >= 0x2000 && c <= 0x206F) return 0;
    if (c >= 0x20A0 && c <= 0x20CF) return 0;
    if (c >= 0x2190 && c <= 0x245F) return 0;
    if (c >= 0x2500 && c <= 0x2BFF) return 0;
    if (c >= 0x2E00 && c <= 0x2E7F) return 0;
    if ((c >= 0x3000 && c <= 0x3004) || (c >= 0x3008 && c <= 0x3020) || c == 0x3030 || (c >= 0x303D && c <= 0x303F)) return 0;
    if ((c >= 0xFE10 && c <= 0xFE1F) || (c >= 0xFE30 && c <= 0xFE6F)) return 0;
    if ((c >= 0xFF00 && c <= 0xFF0F) || (c >= 0xFF1A && c <= 0xFF20) || (c >= 0xFF3B && c <= 0xFF40) || (c >= 0xFF5B && c <= 0xFF65) ||
        (c >= 0xFFE0 && c <= 0xFFEF))
        return 0;
    return 1;
}
static uint32_t cp_lower(uint32_t c) {
    if (c < 0x80) return (uint32_t)tolower((int)c);
    if (c >= 0xC0 && c <= 0xDE && c != 0xD7) return c + 0x20;
    if (c >= 0x100 && c <= 0x137) return (c & 1) ? c : c + 1;
    if (c >= 0x139 && c <= 0x148) return (c & 1) ? c + 1 : c;
    if (c >= 0x14A && c <= 0x177) return (c & 1) ? c : c + 1;
    if (c == 0x178) return 0xFF;
    if (c >= 0x179 && c <= 0x17E) return (c & 1) ? c + 1 : c;
    if (c >= 0x391 && c <= 0x3A9 && c != 0x3A2) return c + 0x20;
    if (c >= 0x410 && c <= 0x42F) return c + 0x20;
    if (c >= 0x400 && c <= 0x40F) return c + 0x50;
    if (c >= 0xFF21 && c <= 0xFF3A) return c + 0x20;
    return c;
}
static void tokenize(const char *text, tok_cb cb, void *ctx) {
    const unsigned char *t = (const unsigned char *)text;
    size_t n = strlen(text), i = 0;
    char buf[192];
    while (i < n) {
        uint32_t cp;
        size_t adv = utf8_decode(t + i, n - i, &cp);
        if (!cp_is_alnum(cp)) { i += adv; continue; }
        size_t s = i, out = 0;
        int too_long = 0;
        while (i < n) {
            adv = utf8_decode(t + i, n - i, &cp);
            if (!cp_is_alnum(cp)) break;
            if (out + 4 < sizeof(buf)) out += utf8_encode(cp_lower(cp), buf + out); else too_long = 1;
            i += adv;
        }
        if (i - s >= 40 || too_long) continue; /* RemoveLongFilter::limit(40): keep len < 40 */
        cb(ctx, buf, out);
    }
}

typedef struct {
    orc_bm25_index *ix;
    uint32_t field, doc;
} add_ctx;

static void add_token(void *c, const char *tok, size_t len) {
    add_ctx *a = (add_ctx *)c;
    orc_bm25_index *ix = a->ix;
    int64_t t = intern_term(ix, a->field, tok, len);
    term_entry *e = &ix->terms[t];
    if (e->n && e->p[e->n - 1].doc == a->doc) {
        e->p[e->n - 1].tf++;
    } else {
        if (e->n == e->cap) {
            e->cap = e->cap ? e->cap * 2 : 4;
            e->p = (posting *)realloc(e->p, sizeof(posting) * e->cap);
        }
        e->p[e->n].doc = a->doc;
        e->p[e->n].tf = 1;
        e->n++;
    }
    ix->doc_len[(size_t)a->field * ix->cap_docs + a->doc]++;
    ix->total_tokens[a->field]++;
}

/* Start a new document (row).  Mirrors ffi_index_multi_column_docs(path, row_id,
 * column_names, column_docs), TantivyIndexStore.cpp:742.  Returns the doc ordinal. */
uint32_t orc_bm25_new_doc(orc_bm25_index *ix, uint64_t row_id) {
    if (ix->n_docs == ix->cap_docs) {
        uint32_t nc = ix->cap_docs ? ix->cap_docs * 2 : 1024;
        uint32_t *dl = (uint32_t *)calloc((size_t)ix->n_fields * nc, sizeof(uint32_t));
        for (uint32_t f = 0; f < ix->n_fields; f++)
            if (ix->cap_docs) memcpy(dl + (size_t)f * nc, ix->doc_len + (size_t)f * ix->cap_docs, sizeof(uint32_t) * ix->n_docs);
        free(ix->doc_len);
        ix->doc_len = dl;
        ix->row_id = (uint64_t *)realloc(ix->row_id, sizeof(uint64_t) * nc);
        ix->cap_docs = nc;
    }
    ix->row_id[ix->n_docs] = row_id;
    return ix->n_docs++;
}

/* Add text to a field of the latest document; call several times for
 * Array(String) columns (multi-valued field: tokens and field norm accumulate). */
void orc_bm25_add_text(orc_bm25_index *ix, uint32_t field, const char *text) {
    add_ctx a = {ix, field, ix->n_docs - 1};
    tokenize(text, add_token, &a);
}

uint64_t orc_bm25_total_docs(const orc_bm25_index *ix) { return ix->n_docs; }
uint64_t orc_bm25_total_tokens(const orc_bm25_index *ix, uint32_t field) { return ix->total_tokens[field]; }
uint64_t orc_bm25_doc_freq(const orc_bm25_index *ix, uint32_t field, const char *term) {
    int64_t t = find_term(ix, field, term, strlen(term));
    return t < 0 ? 0 : ix->terms[t].n;
}

typedef struct {
    char toks[64][48];
    int n;
} qtoks;
static void q_token(void *c, const char *tok, size_t len) {
    qtoks *q = (qtoks *)c;
    if (q->n >= 64) return;
    for (int i = 0; i < q->n; i++)
        if (strlen(q->toks[i]) == len && memcmp(q->toks[i], tok, len) == 0) return; /* distinct terms */
    memcpy(q->toks[q->n], tok, len);
    q->toks[q->n][len] = 0;
    q->n++;
}

/* Tokenise a query sentence into distinct lowercase terms (<= 64); returns count.
 * out: 64 x 48 char buffer. */
int orc_bm25_query_terms(const char *sentence, char *out) {
    qtoks q;
    q.n = 0;
    tokenize(sentence, q_token, &q);
    memcpy(out, q.toks, sizeof(q.toks));
    return q.n;
}

/* ffi_bm25_search(path, sentence, column_names, topk, alive_bitmap, use_filter,
 * enable_nlq(false), operator_or, statistics), TantivyIndexStore.cpp:908-917/:939-948.
 *   fields[n_fields_q]: field ordinals searched
 *   alive: u8 LSB-first bitmap over row ids (bit = 1 alive), used iff use_filter
 *   stat_total_docs / stat_total_tokens[field] / stat_doc_freq[qfield*64 + term]:
 *     table-wide statistics, used iff stat_total_docs > 0, else this part's own.
 * Output: row ids + scores, score descending, ties -> smaller doc.  Returns count. */
uint32_t orc_bm25_search(const orc_bm25_index *ix, const char *sentence, const uint32_t *fields, uint32_t n_fields_q,
                         uint32_t topk, const uint8_t *alive, int use_filter, int operator_or,
                         uint64_t stat_total_docs, const uint64_t *stat_total_tokens, const uint64_t *stat_doc_freq,
                         uint64_t *out_row, float *out_score) {
    fieldnorm_init();
    qtoks q;
    q.n = 0;
    tokenize(sentence, q_token, &q);
    if (q.n == 0 || ix->n_docs == 0 || topk == 0) return 0;
    float *score = (float *)calloc(ix->n_docs, sizeof(float));
    /* tantivy's QueryParser builds AND over TERMS of (OR over the default fields): a document matches when every term
     * is found in at least one of the searched fields */
    uint64_t *hits = (uint64_t *)calloc(ix->n_docs, sizeof(uint64_t));
    const uint64_t all_terms = q.n >= 64 ? ~0ull : ((1ull << q.n) - 1);
    for (uint32_t fq = 0; fq < n_fields_q; fq++) {
        uint32_t f = fields[fq];
        uint64_t N = stat_total_docs ? stat_total_docs : ix->n_docs;
        uint64_t T = stat_total_docs ? stat_total_tokens[f] : ix->total_tokens[f];
        float avgdl = (float)T / (float)N;
        float cache[256];
        for (int c = 0; c < 256; c++)
            cache[c] = BM25_K1 * (1.0f - BM25_B + BM25_B * (float)g_fieldnorm_table[c] / avgdl);
        for (int t = 0; t < q.n; t++) {
            int64_t ti = find_term(ix, f, q.toks[t], strlen(q.toks[t]));
            uint64_t n = stat_total_docs ? stat_doc_freq[fq * 64 + t] : (ti < 0 ? 0 : ix->terms[ti].n);
            if (ti < 0) continue;
            float x = ((float)(N - n) + 0.5f) / ((float)n + 0.5f);
            float idf = logf(1.0f + x);
            float weight = idf * (1.0f + BM25_K1);
            const term_entry *e = &ix->terms[ti];
            for (uint32_t i = 0; i < e->n; i++) {
                uint32_t doc = e->p[i].doc;
                float tf = (float)e->p[i].tf;
                int code = orc_bm25_fieldnorm_to_id(ix->doc_len[(size_t)f * ix->cap_docs + doc]);
                score[doc] += weight * (tf / (tf + cache[code]));
                hits[doc] |= 1ull << t;
            }
        }
    }
    /* top-k: score desc, doc asc */
    uint32_t cnt = 0;
    for (uint32_t doc = 0; doc < ix->n_docs; doc++) {
        if (!hits[doc]) continue;
        if (!operator_or && hits[doc] != all_terms) continue;
        uint64_t rid = ix->row_id[doc];
        if (use_filter && alive && !((alive[rid >> 3] >> (rid & 7)) & 1)) continue;
        float s = score[doc];
        if (cnt == topk && !(s > out_score[topk - 1])) continue;
        uint32_t j = cnt < topk ? cnt++ : topk - 1;
        while (j > 0 && out_score[j - 1] < s) {
            out_score[j] = out_score[j - 1];
            out_row[j] = out_row[j - 1];
            j--;
        }
        out_score[j] = s;
        out_row[j] = rid;
    }
    free(score);
    free(hits);
    return cnt;
}

/* Export postings of one (field, term) for the GPU index builder used in tests:
 * returns df, fills doc ordinals / tfs up to cap. */
uint32_t orc_bm25_postings(const orc_bm25_index *ix, uint32_t field, const char *term, uint32_t *docs, uint32_t *tfs,
                           uint32_t cap) {
    int64_t t = find_term(ix, field, term, strlen(term));
    if (t < 0) return 0;
    const term_entry *e = &ix->terms[t];
    for (uint32_t i = 0; i < e->n && i < cap; i++) {
        docs[i] = e->p[i].doc;
        tfs[i] = e->p[i].tf;
    }
    return e->n;
}

uint32_t orc_bm25_doc_len(const orc_bm25_index *ix, uint32_t field, uint32_t doc) {
    return ix->doc_len[(size_t)field * ix->cap_docs + doc];
}
