/* TEST INFRASTRUCTURE ONLY - the CPU oracle for verticut_b200.
 *
 * A plain-C restatement of the reference's MIH / linear-scan hot path
 * (tu-dresden/verticut), used ONLY by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker.  The product path never links,
 * imports or executes it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors (SURVEY.md
 * section 4), so this restatement is pinned against the reference ITSELF:
 * oracle/_ref/libverticut_ref.so is built from the reference's unmodified
 * sources (oracle/Makefile, oracle/ref_driver.cc) and tests/test_oracle_vs_ref.py
 * requires, on seeded inputs, that
 *   - VO_ORDER_REFERENCE mode reproduces the reference's output exactly
 *     (same ids, same distances, same order, same radius, same probe counts),
 *   - the canonical mode has the same distance multiset and tie-class
 *     membership (contract P3 of SURVEY.md section 8(c)).
 * tests/golden/ holds vectors generated from oracle/_ref by
 * tests/golden/make_golden.py for the box where /root/reference is absent.
 *
 * Every function cites the reference lines it follows.  Deviations from the
 * shipped reference (all listed in DESIGN.md):
 *   D1 binaryToInt is unsigned (the reference sign-extends substrings shorter
 *      than 4 bytes, Pilaf/image_tools.h:12-18; the _ref build uses
 *      -funsigned-char to the same effect).
 *   D2 the stop rule can use the table count m instead of the hard-coded 4
 *      (search_worker.cc:204) and a strict bound (VO_STOP_STRICT_M).
 *   D3 canonical order: results ascending by (dist, id) instead of the
 *      libstdc++-heap-dependent tie order (VO_ORDER_CANONICAL).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VO_FOUND 0      /* base_proxy.h:10 PROXY_FOUND */
#define VO_NOT_FOUND 1  /* base_proxy.h:11 PROXY_NOT_FOUND */

#define VO_ORDER_CANONICAL 0 /* ascending (dist,id); k smallest packed keys */
#define VO_ORDER_REFERENCE 1 /* reference heap semantics, descending dist, libstdc++ tie order */

#define VO_STOP_REF4 0      /* full && top.dist <= (r+1)*4      search_worker.cc:201-205 */
#define VO_STOP_STRICT_M 1  /* full && d_k <= m*(r+1) - 1        SURVEY.md finding 7 */
#define VO_STOP_REF_M 2     /* full && top.dist <= (r+1)*m       finding 6 only */
#define VO_STOP_TABLE_STRICT 3 /* checked after EVERY table t of radius r: full && d_k <= m*r + t.  Codes not yet found
                                  then have substring distance >= r+1 in tables 0..t and >= r in the others, i.e. full
                                  distance >= m*r + t + 1 > d_k: the canonical top-k is final.  Same answers as
                                  VO_STOP_STRICT_M with fewer probes (canonical order only). */

#define VO_APPROX_FACTOR 20 /* search_worker.h:14 APPROXIMATE_FACTOR */

/* ---- a1: Pilaf/image_tools.h:12-18 binaryToInt, unsigned (D1) ---- */
uint32_t vo_binary_to_int(const uint8_t* p, int len) {
  uint32_t result = p[len - 1];
  for (int i = len - 2; i >= 0; i--) result = (result << 8) | p[i];
  return result;
}

/* ---- a2: Pilaf/image_tools.h:21-33 compute_hamming_dist (32-bit words) ---- */
int vo_hamming(const uint8_t* a, const uint8_t* b, int nbytes) {
  int dist = 0;
  for (int i = 0; i + 4 <= nbytes; i += 4) {
    uint32_t x, y;
    memcpy(&x, a + i, 4);
    memcpy(&y, b + i, 4);
    dist += __builtin_popcount(x ^ y);
  }
  return dist;
}

/* ---- a13: packed candidate word, search_worker.cc:12-13,254-256 ---- */
static inline uint64_t pack(uint32_t dist, uint32_t id) { return ((uint64_t)dist << 32) | id; }

/* ---- a10: src/bitmap.cc:22-38 ImageBitmap::get_idx / set_idx / reset_idx ---- */
int vo_bitmap_get(const uint32_t* data, uint64_t bit) { return (data[bit / 32] >> (bit % 32)) & 1u; }
void vo_bitmap_set(uint32_t* data, uint64_t bit) { data[bit / 32] |= (1u << (bit % 32)); }
void vo_bitmap_reset(uint32_t* data, uint64_t bit) { data[bit / 32] &= ~(1u << (bit % 32)); }

/* =====================================================================
 * a9: index build.  build_hash_tables.cc:25-73: id = ordinal in the code
 * stream; table t keys on bytes [t*sub, (t+1)*sub); a bucket lists its
 * members in insertion (= ascending id) order.  Stored as one stable
 * key-sorted permutation per table.
 * ===================================================================== */
typedef struct {
  const uint8_t* codes; /* borrowed, n * nbytes */
  uint64_t n;
  int nbytes, m, sub; /* sub = bytes per substring */
  uint32_t first_id;
  uint32_t** skeys; /* [m][n] sorted keys */
  uint32_t** order; /* [m][n] ordinal of the code at each sorted position */
} vo_index;

static void radix_sort_pairs(uint32_t* keys, uint32_t* vals, uint64_t n, int key_bytes) {
  /* stable LSD radix sort, 8-bit digits; result left in keys/vals */
  uint32_t* kb = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t* vb = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t *ks = keys, *vs = vals, *kd = kb, *vd = vb;
  for (int pass = 0; pass < key_bytes; ++pass) {
    uint64_t cnt[257];
    memset(cnt, 0, sizeof cnt);
    int sh = pass * 8;
    for (uint64_t i = 0; i < n; ++i) cnt[((ks[i] >> sh) & 0xff) + 1]++;
    for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
    for (uint64_t i = 0; i < n; ++i) { /* stable: ascending ordinal within a key */
      uint64_t dst = cnt[(ks[i] >> sh) & 0xff]++;
      kd[dst] = ks[i];
      vd[dst] = vs[i];
    }
    uint32_t* t = ks; ks = kd; kd = t;
    t = vs; vs = vd; vd = t;
  }
  if (ks != keys) {
    memcpy(keys, ks, n * sizeof(uint32_t));
    memcpy(vals, vs, n * sizeof(uint32_t));
  }
  free(kb);
  free(vb);
}

vo_index* vo_index_create(const uint8_t* codes, uint64_t n, int nbytes, int m, uint32_t first_id) {
  if (m <= 0 || nbytes % m != 0 || nbytes / m > 4 || nbytes % 4 != 0) return 0; /* search_worker.cc:75; uint32 index */
  vo_index* ix = (vo_index*)calloc(1, sizeof(vo_index));
  ix->codes = codes; ix->n = n; ix->nbytes = nbytes; ix->m = m; ix->sub = nbytes / m; ix->first_id = first_id;
  ix->skeys = (uint32_t**)calloc(m, sizeof(uint32_t*));
  ix->order = (uint32_t**)calloc(m, sizeof(uint32_t*));
  for (int t = 0; t < m; ++t) {
    uint32_t* k = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t* v = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < n; ++i) {
      k[i] = vo_binary_to_int(codes + i * nbytes + t * ix->sub, ix->sub); /* build_hash_tables.cc:38,45 */
      v[i] = (uint32_t)i;
    }
    radix_sort_pairs(k, v, n, ix->sub);
    ix->skeys[t] = k;
    ix->order[t] = v;
  }
  return ix;
}

void vo_index_destroy(vo_index* ix) {
  if (!ix) return;
  for (int t = 0; t < ix->m; ++t) { free(ix->skeys[t]); free(ix->order[t]); }
  free(ix->skeys); free(ix->order); free(ix);
}

static uint64_t lower_bound_u32(const uint32_t* a, uint64_t n, uint64_t key) { /* first i with a[i] >= key */
  uint64_t lo = 0, hi = n;
  while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}

/* BaseProxy::get(HashIndex{table,index}, Image_List) - base_proxy.h:18; members in stored order */
int vo_bucket_get(const vo_index* ix, uint32_t table, uint32_t key, uint32_t* ids, uint32_t cap, uint32_t* n_out) {
  if (table >= (uint32_t)ix->m) { *n_out = 0; return VO_NOT_FOUND; }
  uint64_t b = lower_bound_u32(ix->skeys[table], ix->n, key);
  uint64_t e = lower_bound_u32(ix->skeys[table], ix->n, (uint64_t)key + 1);
  *n_out = (uint32_t)(e - b);
  for (uint64_t i = b; i < e && i - b < cap; ++i) ids[i - b] = ix->first_id + ix->order[table][i];
  return e > b ? VO_FOUND : VO_NOT_FOUND;
}

/* generate_bitmap.cc:54-58,105-119: occupancy bitmap of table t (bit = bucket index non-empty) */
void vo_occupancy_bitmap(const vo_index* ix, uint32_t table, uint32_t* data /* 2^(8*sub)/32 words, zeroed */) {
  for (uint64_t i = 0; i < ix->n; ++i) vo_bitmap_set(data, ix->skeys[table][i]);
}

/* =====================================================================
 * libstdc++'s std::priority_queue<T> (max-heap on dist only), restated so
 * that VO_ORDER_REFERENCE reproduces the reference's tie order:
 * operator< compares dist only (search_worker.cc:15-17, linear_search.cc:34-37).
 * push  = vector::push_back + std::push_heap  (__push_heap: sift the hole up
 *         while parent < value)
 * pop   = std::pop_heap + pop_back (__pop_heap: move last out, __adjust_heap
 *         from the root: walk the larger child down to a leaf - on equal
 *         children take the RIGHT one unless right < left - then __push_heap)
 * ===================================================================== */
typedef struct { uint32_t id, dist; } vo_item;
typedef struct { vo_item* a; size_t n, cap; } vo_heap;

static void heap_push(vo_heap* h, vo_item v) {
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->a = (vo_item*)realloc(h->a, h->cap * sizeof(vo_item)); }
  size_t hole = h->n++;
  while (hole > 0) {
    size_t parent = (hole - 1) / 2;
    if (!(h->a[parent].dist < v.dist)) break;
    h->a[hole] = h->a[parent];
    hole = parent;
  }
  h->a[hole] = v;
}

static vo_item heap_pop(vo_heap* h) {
  vo_item top = h->a[0];
  size_t len = --h->n; /* elements [0,len) remain; value = old a[len] */
  if (len == 0) return top;
  vo_item v = h->a[len];
  size_t hole = 0, child = 0;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (h->a[child].dist < h->a[child - 1].dist) child--;
    h->a[hole] = h->a[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h->a[hole] = h->a[child - 1];
    hole = child - 1;
  }
  while (hole > 0) { /* __push_heap(first, hole, 0, value) */
    size_t parent = (hole - 1) / 2;
    if (!(h->a[parent].dist < v.dist)) break;
    h->a[hole] = h->a[parent];
    hole = parent;
  }
  h->a[hole] = v;
  return top;
}

/* canonical top-k: sorted ascending array of at most k packed keys */
typedef struct { uint64_t* a; size_t n, k; } vo_topk;
static void topk_offer(vo_topk* t, uint64_t key) {
  if (t->k == 0) return;
  if (t->n == t->k && key >= t->a[t->n - 1]) return;
  size_t i = t->n < t->k ? t->n++ : t->n - 1;
  while (i > 0 && t->a[i - 1] > key) { t->a[i] = t->a[i - 1]; --i; }
  t->a[i] = key;
}

/* =====================================================================
 * a8: linear_search.cc:39-64 search_K_nearest_neighbors(int k)
 * ===================================================================== */
/* Canonical: the k smallest (dist<<32 | id), ascending.  Returns the count (min(k, n)). */
uint32_t vo_linear_search(const uint8_t* codes, uint64_t n, int nbytes, uint32_t first_id, const uint8_t* query,
                          uint32_t k, uint32_t* out_ids, uint32_t* out_dists) {
  vo_topk t = {(uint64_t*)malloc((k ? k : 1) * sizeof(uint64_t)), 0, k};
  for (uint64_t i = 0; i < n; ++i) /* linear_search.cc:44-50 */
    topk_offer(&t, pack((uint32_t)vo_hamming(codes + i * nbytes, query, nbytes), first_id + (uint32_t)i));
  for (size_t i = 0; i < t.n; ++i) { out_ids[i] = (uint32_t)t.a[i]; out_dists[i] = (uint32_t)(t.a[i] >> 32); }
  uint32_t cnt = (uint32_t)t.n;
  free(t.a);
  return cnt;
}

/* Reference order: ids ascending into a max-heap with strict-less replacement (linear_search.cc:51-56),
   popped to the output in descending distance (:59-63). */
uint32_t vo_linear_search_ref(const uint8_t* codes, uint64_t n, int nbytes, uint32_t first_id, const uint8_t* query,
                              uint32_t k, uint32_t* out_ids, uint32_t* out_dists) {
  vo_heap h = {0, 0, 0};
  for (uint64_t i = 0; i < n; ++i) {
    vo_item it = {first_id + (uint32_t)i, (uint32_t)vo_hamming(codes + i * nbytes, query, nbytes)};
    if (h.n < k) heap_push(&h, it);
    else if (h.n && h.a[0].dist > it.dist) { heap_pop(&h); heap_push(&h, it); }
  }
  uint32_t cnt = 0;
  while (h.n) { vo_item it = heap_pop(&h); out_ids[cnt] = it.id; out_dists[cnt] = it.dist; ++cnt; }
  free(h.a);
  return cnt;
}

/* =====================================================================
 * a3-a7: SearchWorker::find and friends, search_worker.cc:65-264
 * ===================================================================== */
typedef struct {
  uint32_t radius;       /* last radius searched (search_worker.cc:217; A.7) */
  uint64_t probes;       /* n_sub_reads_ summed over tables (search_worker.cc:245) */
  uint64_t candidates;   /* members of probed buckets, duplicates included */
  uint64_t unique;       /* distinct ids seen (knn_found_.size(), search_worker.cc:190) */
} vo_stats;

typedef struct {
  const vo_index* ix;
  const uint8_t* query;
  int sbits;
  uint32_t table;
  uint64_t* cand; size_t ncand, capcand; /* kn_candidates of the current radius, all tables in rank order */
  vo_stats* st;
} vo_enum;

/* leaf of enumerate_entry, search_worker.cc:233-258 (occupancy bitmap disabled as shipped, :59-62) */
static void probe_bucket(vo_enum* e, uint32_t curr) {
  const vo_index* ix = e->ix;
  e->st->probes++;
  uint64_t b = lower_bound_u32(ix->skeys[e->table], ix->n, curr);
  for (uint64_t i = b; i < ix->n && ix->skeys[e->table][i] == curr; ++i) {
    uint32_t ord = ix->order[e->table][i];
    uint32_t dist = (uint32_t)vo_hamming(ix->codes + (uint64_t)ord * ix->nbytes, e->query, ix->nbytes);
    if (e->ncand == e->capcand) {
      e->capcand = e->capcand ? e->capcand * 2 : 8192;
      e->cand = (uint64_t*)realloc(e->cand, e->capcand * sizeof(uint64_t));
    }
    e->cand[e->ncand++] = pack(dist, ix->first_id + ord); /* search_worker.cc:254-256 */
    e->st->candidates++;
  }
}

/* search_worker.cc:230-264 enumerate_entry: every index at Hamming distance exactly rr, flip-low-bit-first DFS */
static void enumerate_entry(vo_enum* e, uint32_t curr, int len, int rr) {
  if (rr == 0) { probe_bucket(e, curr); return; }
  enumerate_entry(e, curr ^ (1u << len), len + 1, rr - 1); /* :260 */
  if (e->sbits - len > rr) enumerate_entry(e, curr, len + 1, rr); /* :261-262 */
}

/* One query.  flags: order (VO_ORDER_*), stop (VO_STOP_*), approximate (search_worker.cc:93-157),
 * max_radius >= 0 -> fixed-radius mode: search radii 0..max_radius, no stop rule (config C5).
 * Output: out_ids/out_dists [k]; canonical = ascending (dist,id), reference = descending dist.
 * Returns the number of results. */
uint32_t vo_mih_search(const vo_index* ix, const uint8_t* query, uint32_t k, int order, int stop, int approximate,
                       int max_radius, uint32_t* out_ids, uint32_t* out_dists, vo_stats* st_out) {
  vo_stats st = {0, 0, 0, 0};
  int sbits = ix->sub * 8;
  uint64_t nseen_words = (ix->n + 63) / 64 + 1;
  uint64_t* seen = (uint64_t*)calloc(nseen_words, sizeof(uint64_t)); /* knn_found_ (search_worker.h:40) */
  vo_enum e = {ix, query, sbits, 0, 0, 0, 0, &st};
  size_t heap_cap = approximate ? (size_t)k * VO_APPROX_FACTOR : (size_t)k; /* :126 vs :192 */
  vo_heap h = {0, 0, 0};
  vo_topk tk = {(uint64_t*)malloc((k ? k : 1) * sizeof(uint64_t)), 0, k};
  uint32_t search_index[64];
  for (int t = 0; t < ix->m; ++t) search_index[t] = vo_binary_to_int(query + t * ix->sub, ix->sub); /* :165-167 */

  int radius = 0, is_stop = 0;
  while (!is_stop && radius <= sbits) { /* :170 */
    e.ncand = 0;
    for (int t = 0; t < ix->m; ++t) { /* ranks = tables; gather_vectors concatenates in rank order (mpi_coordinator.cc:50-59) */
      e.table = (uint32_t)t;
      enumerate_entry(&e, search_index[t], 0, radius); /* :222-227 */
      if (stop == VO_STOP_TABLE_STRICT && order == VO_ORDER_CANONICAL && !approximate && max_radius < 0) {
        /* table-granular variant: digest this table's candidates now and test the stop rule */
        for (size_t i = 0; i < e.ncand; ++i) {
          uint32_t id = (uint32_t)e.cand[i];
          uint64_t ord = id - ix->first_id;
          if (seen[ord / 64] >> (ord % 64) & 1) continue;
          seen[ord / 64] |= 1ull << (ord % 64);
          st.unique++;
          topk_offer(&tk, e.cand[i]);
        }
        e.ncand = 0;
        if (tk.n == tk.k && tk.k > 0 && (uint32_t)(tk.a[tk.n - 1] >> 32) <= (uint32_t)(ix->m * radius + t)) { is_stop = 1; break; }
      }
    }
    if (is_stop) { radius += 1; break; }
    for (size_t i = 0; i < e.ncand; ++i) { /* master loop :179-199 */
      uint32_t id = (uint32_t)e.cand[i], dist = (uint32_t)(e.cand[i] >> 32);
      uint64_t ord = id - ix->first_id;
      if (seen[ord / 64] >> (ord % 64) & 1) continue; /* :183-184 */
      seen[ord / 64] |= 1ull << (ord % 64);            /* :190 */
      st.unique++;
      if (order == VO_ORDER_REFERENCE) {
        vo_item it = {id, dist};
        if (h.n < heap_cap) heap_push(&h, it);                                   /* :192-193 */
        else if (h.n && h.a[0].dist > dist) { heap_pop(&h); heap_push(&h, it); } /* :194-197 */
      } else {
        topk_offer(&tk, e.cand[i]);
      }
    }
    radius += 1; /* :201 */
    if (max_radius >= 0) {
      if (radius > max_radius) is_stop = 1;
    } else if (approximate) {
      /* :136-137 heap full; canonical: as many distinct candidates as the heap would hold */
      if (order == VO_ORDER_REFERENCE ? (h.n == heap_cap) : (st.unique >= heap_cap)) is_stop = 1;
    } else if (order == VO_ORDER_REFERENCE) {
      uint32_t mult = stop == VO_STOP_REF4 ? 4u : (uint32_t)ix->m;
      if (h.n == heap_cap && heap_cap > 0) {
        if (stop == VO_STOP_STRICT_M ? (h.a[0].dist + 1 <= (uint32_t)radius * mult) : (h.a[0].dist <= (uint32_t)radius * mult))
          is_stop = 1; /* :204-205 */
      }
    } else {
      if (tk.n == tk.k && tk.k > 0) {
        uint32_t dk = (uint32_t)(tk.a[tk.n - 1] >> 32);
        uint32_t mult = stop == VO_STOP_REF4 ? 4u : (uint32_t)ix->m;
        if (stop == VO_STOP_STRICT_M || stop == VO_STOP_TABLE_STRICT ? (dk + 1 <= (uint32_t)radius * mult) : (dk <= (uint32_t)radius * mult)) is_stop = 1;
      }
    }
  }
  st.radius = (uint32_t)(radius - 1); /* :217 */

  uint32_t cnt = 0;
  if (order == VO_ORDER_REFERENCE) {
    size_t beg = (approximate && h.n > k) ? h.n - k : 0; /* :143 */
    size_t i = 0;
    while (h.n) { /* :145-155 / :210-216 */
      vo_item it = heap_pop(&h);
      if (i >= beg && cnt < k) { out_ids[cnt] = it.id; out_dists[cnt] = it.dist; ++cnt; }
      ++i;
    }
  } else {
    for (size_t i = 0; i < tk.n; ++i) { out_ids[cnt] = (uint32_t)tk.a[i]; out_dists[cnt] = (uint32_t)(tk.a[i] >> 32); ++cnt; }
  }
  if (st_out) *st_out = st;
  free(seen); free(e.cand); free(h.a); free(tk.a);
  return cnt;
}

/* a12/K7: merge of per-shard local top-k lists (replaces gather_vectors + master heap,
 * mpi_coordinator.cc:34-69 / search_worker.cc:179-199, for id-disjoint shards): the k smallest
 * packed keys of the union, ascending.  lists = n_lists * k keys, unused slots = UINT64_MAX. */
uint32_t vo_merge_topk(const uint64_t* lists, uint32_t n_lists, uint32_t k, uint64_t* out) {
  vo_topk t = {out, 0, k};
  for (uint64_t i = 0; i < (uint64_t)n_lists * k; ++i)
    if (lists[i] != UINT64_MAX) topk_offer(&t, lists[i]);
  return (uint32_t)t.n;
}

/* synthetic data generator shared with the product's on-device generator:
 * word w of code id = splitmix64 finaliser of (seed, id, w); little-endian 64-bit words. */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
uint64_t vo_synth_word(uint64_t seed, uint64_t id, uint32_t word) {
  return splitmix64(splitmix64(seed ^ (0xD1B54A32D192ED03ull * (word + 1))) + id);
}
void vo_synth_codes(uint64_t seed, uint64_t first_id, uint64_t n, int nbytes, uint8_t* out) {
  int words = nbytes / 8;
  for (uint64_t i = 0; i < n; ++i)
    for (int w = 0; w < words; ++w) {
      uint64_t v = vo_synth_word(seed, first_id + i, (uint32_t)w);
      memcpy(out + i * nbytes + (size_t)w * 8, &v, 8);
    }
}

/* Streaming form of a8 (linear_search.cc:39-64) over the SYNTHETIC database, for checks at sizes whose codes no host can
 * hold: codes i0 <= i < i1 of a shard whose code i has id first_id + i * stride are generated on the fly and offered to one
 * canonical top-k per query.  With m > 0 and max_radius >= 0 only those codes count that the fixed-radius MIH search reaches:
 * some table t holds them in a bucket within substring distance max_radius of the query's substring t - exactly the union of
 * the buckets SearchWorker::search_R_neighbors / enumerate_entry visit for radii 0 .. max_radius (search_worker.cc:222-264;
 * pinned against vo_mih_search in tests/test_oracle_vs_ref.py).  out_keys: [nq][k] packed words, UINT64_MAX padded. */
void vo_scan_synth(uint64_t seed, uint64_t first_id, uint64_t stride, uint64_t i0, uint64_t i1, int nbytes, const uint8_t* queries,
                   uint32_t nq, uint32_t k, int m, int max_radius, uint64_t* out_keys) {
  const int words = nbytes / 8, sub = m > 0 ? nbytes / m : 0;
  uint64_t base[8];
  uint8_t code[64];
  vo_topk* tk = (vo_topk*)malloc(sizeof(vo_topk) * (nq ? nq : 1));
  for (uint32_t q = 0; q < nq; ++q) { tk[q].a = out_keys + (size_t)q * k; tk[q].n = 0; tk[q].k = k; }
  for (int w = 0; w < words; ++w) base[w] = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(w + 1)));   /* vo_synth_word, hoisted */
  for (uint64_t i = i0; i < i1; ++i) {
    const uint64_t id = first_id + i * stride;
    for (int w = 0; w < words; ++w) { const uint64_t v = splitmix64(base[w] + id); memcpy(code + 8 * w, &v, 8); }
    for (uint32_t q = 0; q < nq; ++q) {
      const uint8_t* qc = queries + (size_t)q * nbytes;
      if (m > 0 && max_radius >= 0) {
        int reached = 0;
        for (int t = 0; t < m && !reached; ++t) {   /* substrings are 1, 2 or 4 bytes: compute_hamming_dist counts whole 32-bit words only */
          int sd = 0;
          for (int b = 0; b < sub; ++b) sd += __builtin_popcount((unsigned)(code[t * sub + b] ^ qc[t * sub + b]));
          reached = sd <= max_radius;
        }
        if (!reached) continue;
      }
      topk_offer(&tk[q], pack((uint32_t)vo_hamming(code, qc, nbytes), (uint32_t)id));
    }
  }
  for (uint32_t q = 0; q < nq; ++q)
    for (size_t j = tk[q].n; j < k; ++j) out_keys[(size_t)q * k + j] = UINT64_MAX;
  free(tk);
}
