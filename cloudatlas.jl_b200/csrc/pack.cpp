#include "pack.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <atomic>
#include <thread>
#include <unordered_set>

#include "common.h"

namespace ca {

namespace {

constexpr int kSketch = 8;

inline uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

struct Row {
  std::vector<uint64_t> keys;  // sorted pair ids a*nin+b
  std::vector<double> vals;
  uint64_t sketch[kSketch];
};

size_t union_size(const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
  size_t i = 0, j = 0, n = 0;
  while (i < a.size() && j < b.size()) {
    if (a[i] == b[j]) { ++i; ++j; }
    else if (a[i] < b[j]) ++i;
    else ++j;
    ++n;
  }
  return n + (a.size() - i) + (b.size() - j);
}

std::vector<uint64_t> union_of(const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
  std::vector<uint64_t> out;
  out.reserve(a.size() + b.size());
  std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
  return out;
}


// Row key sets as the host packer holds them: sorted vectors.
struct HostRowSets final : RowSetOps {
  const std::vector<Row>& rows;
  std::vector<uint64_t> U;
  explicit HostRowSets(const std::vector<Row>& r) : rows(r) {}
  int32_t nrows() const override { return (int32_t)rows.size(); }
  int64_t nkeys(int32_t r) const override { return (int64_t)rows[r].keys.size(); }
  const uint64_t* sketch(int32_t r) const override { return rows[r].sketch; }
  int batch_hint() const override { return 1; }
  void open(int32_t seed) override { U = rows[seed].keys; }
  int64_t open_size() const override { return (int64_t)U.size(); }
  void union_sizes(const int32_t* cand, int n, int64_t* out) override {
    for (int q = 0; q < n; ++q) out[q] = (int64_t)union_size(U, rows[cand[q]].keys);
  }
  void absorb(int32_t r) override { U = union_of(U, rows[r].keys); }
};

}  // namespace

uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
  // 8 bytes at a time (the arrays digested are gigabytes at m ~ 3000), remainder bytewise
  const unsigned char* p = (const unsigned char*)data;
  size_t q = 0;
  for (; q + 8 <= bytes; q += 8) {
    uint64_t w;
    memcpy(&w, p + q, 8);
    h = (h ^ w) * 0x100000001b3ull;
  }
  for (; q < bytes; ++q) h = (h ^ p[q]) * 0x100000001b3ull;
  return h;
}

void pack_digest(const PackedHost& P, uint64_t out[8]) {
  // descriptors are hashed field by field (struct padding is not part of a pack)
  uint64_t h = 0xcbf29ce484222325ull;
  for (const TileDesc& T : P.tiles) {
    const int64_t f[7] = {T.ks_begin, T.nks, T.vals_off, T.rows_off, T.nt, T.nrows, T.npairs};
    h = fnv1a(f, sizeof f, h);
  }
  out[0] = h;
  out[1] = fnv1a(P.pairs.data(), P.pairs.size() * sizeof(uint32_t));
  out[2] = fnv1a(P.vals.data(), P.vals.size() * sizeof(double));
  out[3] = fnv1a(P.rows.data(), P.rows.size() * sizeof(int32_t));
  h = 0xcbf29ce484222325ull;
  for (const WindowDesc& W : P.windows) {
    const int64_t f[9] = {W.ks_begin, W.nks, W.modes_off, W.nmodes, W.tile, W.last, W.rows_off, W.nt, W.vals_off};
    h = fnv1a(f, sizeof f, h);
  }
  out[4] = h;
  out[5] = fnv1a(P.modes.data(), P.modes.size() * sizeof(int32_t));
  h = fnv1a(P.sched.data(), P.sched.size() * sizeof(int32_t));
  h = fnv1a(P.sched_off, sizeof P.sched_off, h);
  h = fnv1a(P.sched_nt2, sizeof P.sched_nt2, h);
  h = fnv1a(P.sched_ks2, sizeof P.sched_ks2, h);
  h = fnv1a(P.sched_ks1, sizeof P.sched_ks1, h);
  const int64_t sl = P.stream_layout;
  out[6] = fnv1a(&sl, sizeof sl, h);
  const int64_t c[9] = {P.m, P.nin, P.symmetric, P.window_modes, P.max_window_modes, P.nterms, P.npairs_distinct,
                        P.padded_fma, P.pair_products};
  out[7] = fnv1a(c, sizeof c);
}

void cluster_rows(RowSetOps& sets, std::vector<std::vector<int32_t>>& clusters, std::vector<int32_t>& empties) {
  const int32_t m = sets.nrows();
  std::vector<int32_t> order(m);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return sets.nkeys(x) > sets.nkeys(y); });
  std::vector<char> assigned(m, 0);
  const int batch = std::max(1, sets.batch_hint());
  std::vector<int32_t> cq;
  std::vector<int64_t> usz;
  for (int32_t seed : order) {
    if (assigned[seed]) continue;
    assigned[seed] = 1;
    if (sets.nkeys(seed) == 0) {
      empties.push_back(seed);
      continue;
    }
    std::vector<int32_t> cl{seed};
    sets.open(seed);
    int64_t usize = sets.nkeys(seed);
    std::vector<std::pair<int, int32_t>> cand;
    const uint64_t* ss = sets.sketch(seed);
    for (int32_t r = 0; r < m; ++r) {
      if (assigned[r] || sets.nkeys(r) == 0) continue;
      const uint64_t* sr = sets.sketch(r);
      int agree = 0;
      for (int h = 0; h < kSketch; ++h) agree += sr[h] == ss[h];
      if (agree >= kSketch / 2) cand.push_back({-agree, r});
    }
    std::sort(cand.begin(), cand.end());
    // union sizes are asked for in batches against the open set; a batch is void as soon as the open set grows
    size_t next = 0;
    while (next < cand.size()) {
      const size_t n = std::min(cand.size() - next, (size_t)batch);
      cq.resize(n);
      usz.resize(n);
      for (size_t q = 0; q < n; ++q) cq[q] = cand[next + q].second;
      sets.union_sizes(cq.data(), (int)n, usz.data());
      size_t q = 0;
      for (; q < n; ++q) {
        const int64_t k = sets.nkeys(cq[q]);
        const int64_t u = usz[q];
        if ((double)u <= 1.2 * (double)std::max(usize, k)) {
          cl.push_back(cq[q]);
          assigned[cq[q]] = 1;
          if (u != usize) {  // the open set grows: the remaining sizes of this batch are stale
            sets.absorb(cq[q]);
            usize = u;
            ++q;
            break;
          }
        }
      }
      next += q;
    }
    std::sort(cl.begin(), cl.end());
    clusters.push_back(std::move(cl));
  }
}

void sort_terms(std::vector<Term>& terms, int32_t m) {
  if (m <= 0) {
    for (const Term& t : terms) m = std::max(m, t.i + 1);
  }
  if (terms.size() < (size_t)1 << 16) {
    std::sort(terms.begin(), terms.end(), [](const Term& x, const Term& y) {
      if (x.i != y.i) return x.i < y.i;
      if (x.a != y.a) return x.a < y.a;
      return x.b < y.b;
    });
    return;
  }
  // Runs of equal output row (assembly, folding and Julia's COO all emit whole rows): found chunk-wise in parallel,
  // then moved as blocks.  Input without such runs degenerates to one run per term and costs a little more than the
  // element-wise scatter it replaces.
  const size_t n = terms.size();
  const size_t nchunks = 64;
  std::vector<std::vector<size_t>> chunk_runs(nchunks);
  parallel_chunks(n, nchunks, [&](size_t b, size_t e, size_t c) {
    std::vector<size_t>& r = chunk_runs[c];
    for (size_t q = b; q < e; ++q)
      if (q == 0 || terms[q].i != terms[q - 1].i) r.push_back(q);
  });
  std::vector<size_t> run_begin;
  for (auto& r : chunk_runs) run_begin.insert(run_begin.end(), r.begin(), r.end());
  run_begin.push_back(n);
  const size_t nruns = run_begin.size() - 1;
  std::vector<size_t> start((size_t)m + 1, 0);
  for (size_t r = 0; r < nruns; ++r) start[(size_t)terms[run_begin[r]].i + 1] += run_begin[r + 1] - run_begin[r];
  for (int32_t i = 0; i < m; ++i) start[(size_t)i + 1] += start[i];
  bool in_place = true;  // already grouped by ascending row: nothing to move
  for (size_t r = 1; r < nruns && in_place; ++r) in_place = terms[run_begin[r]].i > terms[run_begin[r - 1]].i;
  if (!in_place && nruns > n / 8) {  // no runs to speak of: element-wise counting sort
    std::vector<Term> out(n);
    std::vector<size_t> at(start.begin(), start.end() - 1);
    for (const Term& t : terms) out[at[t.i]++] = t;
    terms.swap(out);
  } else if (!in_place) {
    std::vector<size_t> dest(nruns);
    {
      std::vector<size_t> at(start.begin(), start.end() - 1);
      for (size_t r = 0; r < nruns; ++r) {
        dest[r] = at[terms[run_begin[r]].i];
        at[terms[run_begin[r]].i] += run_begin[r + 1] - run_begin[r];
      }
    }
    std::vector<Term> out(n);
    parallel_chunks(nruns, std::min<size_t>(nruns, 1024), [&](size_t b, size_t e, size_t) {
      for (size_t r = b; r < e; ++r)
        std::copy(terms.begin() + (ptrdiff_t)run_begin[r], terms.begin() + (ptrdiff_t)run_begin[r + 1],
                  out.begin() + (ptrdiff_t)dest[r]);
    });
    terms.swap(out);
  }
  const unsigned nthreads = ca::worker_threads();
  std::atomic<int32_t> next{0};
  auto work = [&]() {
    auto less = [](const Term& x, const Term& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; };
    for (int32_t i = next.fetch_add(1); i < m; i = next.fetch_add(1)) {
      auto b = terms.begin() + (ptrdiff_t)start[i], e = terms.begin() + (ptrdiff_t)start[(size_t)i + 1];
      if (!std::is_sorted(b, e, less)) std::sort(b, e, less);  // assembly and folding emit rows already sorted
    }
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
}

std::vector<Term> merge_terms(std::vector<Term> terms, bool symmetric, int32_t m) {
  for (auto& t : terms)
    if (symmetric && t.a > t.b) std::swap(t.a, t.b);
  sort_terms(terms, m);
  std::vector<Term> out;
  out.reserve(terms.size());
  for (size_t r = 0; r < terms.size();) {
    size_t e = r;
    double s = 0;
    while (e < terms.size() && terms[e].i == terms[r].i && terms[e].a == terms[r].a && terms[e].b == terms[r].b)
      s += terms[e++].v;
    if (s != 0.0) out.push_back(Term{terms[r].i, terms[r].a, terms[r].b, s});
    r = e;
  }
  return out;
}

PackedHost pack_terms(std::vector<Term> terms, int32_t m, bool symmetric, int nt_max, int window_modes,
                      int32_t n_inputs) {
  if (m <= 0) fail(CA_ERR_INVALID, "pack_terms: m must be positive");
  const int32_t nin = (n_inputs > 0 ? n_inputs : m) + 1;
  if (nin > 32768) fail(CA_ERR_UNSUPPORTED, "pack_terms: m = %d exceeds the 15-bit mode index", m);
  if (nt_max < 1) nt_max = 1;
  PhaseTimer tm("pack_terms");
  PackedHost P;
  P.m = m;
  P.nin = nin;
  P.symmetric = symmetric;
  P.window_modes = window_modes;

  {
    std::atomic<long long> bad{-1};
    parallel_chunks(terms.size(), 64, [&](size_t b, size_t e, size_t) {
      for (size_t q = b; q < e; ++q) {
        Term& t = terms[q];
        if (t.i < 0 || t.i >= m || t.a < 0 || t.a >= nin || t.b < 0 || t.b >= nin) bad = (long long)q;
        else if (symmetric && t.a > t.b) std::swap(t.a, t.b);
      }
    });
    if (bad >= 0) {
      const Term& t = terms[(size_t)bad.load()];
      fail(CA_ERR_INVALID, "pack_terms: index out of range (i=%d a=%d b=%d, m=%d)", t.i, t.a, t.b, m);
    }
  }
  sort_terms(terms, m);

  tm.lap("sort");
  std::vector<Row> rows(m);
  {
    // terms are sorted by (i, a, b): rows are contiguous ranges, built (merge duplicates, min-hash sketch) on
    // worker threads
    std::vector<size_t> start((size_t)m + 1, 0);
    for (const Term& t : terms) start[(size_t)t.i + 1]++;
    for (int32_t i = 0; i < m; ++i) start[(size_t)i + 1] += start[i];
    std::vector<int64_t> kept(m, 0);
    std::atomic<int32_t> next{0};
    auto work = [&]() {
      for (int32_t i = next.fetch_add(1); i < m; i = next.fetch_add(1)) {
        Row& rw = rows[i];
        size_t r = start[i];
        const size_t end = start[(size_t)i + 1];
        rw.keys.reserve(end - r);
        rw.vals.reserve(end - r);
        while (r < end) {
          size_t e = r;
          double sum = 0;
          while (e < end && terms[e].a == terms[r].a && terms[e].b == terms[r].b) sum += terms[e++].v;
          if (sum != 0.0) {
            rw.keys.push_back((uint64_t)terms[r].a * nin + terms[r].b);
            rw.vals.push_back(sum);
          }
          r = e;
        }
        kept[i] = (int64_t)rw.keys.size();
        for (int h = 0; h < kSketch; ++h) {
          uint64_t best = ~0ull;
          for (uint64_t k : rw.keys) best = std::min(best, mix64(k * 0x100000001b3ull + h * 0x51ed27ull));
          rw.sketch[h] = best;
        }
      }
    };
    const unsigned nthreads = terms.size() < ((size_t)1 << 16) ? 1u : ca::worker_threads();
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    for (int32_t i = 0; i < m; ++i) P.nterms += kept[i];
  }
  {  // distinct (a, b) over the whole operator: one bit per possible pair
    std::vector<uint64_t> seen(((size_t)nin * nin + 63) / 64, 0);
    for (auto& rw : rows)
      for (uint64_t k : rw.keys) seen[k >> 6] |= 1ull << (k & 63);
    int64_t n = 0;
    for (uint64_t w : seen) n += __builtin_popcountll(w);
    P.npairs_distinct = n;
  }

  tm.lap("rows+sketches");
  // ---- cluster rows with near-identical pair sets
  HostRowSets host_sets(rows);
  std::vector<std::vector<int32_t>> clusters;
  std::vector<int32_t> empties;
  cluster_rows(host_sets, clusters, empties);

  nt_max = decide_nt_max(clusters, m, nt_max);
  tm.lap("cluster");
  if (getenv("CA_TIMING") && clusters.size() > 8) {
    int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // rows per cluster: 1-8, 9-16, 17-24, ...
    for (auto& cl : clusters) hist[std::min<size_t>(7, (cl.size() - 1) / 8)]++;
    fprintf(stderr, "[ca timing] pack_terms: %zu row clusters; sizes 1-8: %d, 9-16: %d, 17-24: %d, 25-32: %d, more: %d\n",
            clusters.size(), hist[0], hist[1], hist[2], hist[3], hist[4] + hist[5] + hist[6] + hist[7]);
  }
  // ---- split clusters into tiles of <= 8*nt_max rows, lay out fragments
  // (tiles are laid out independently on worker threads, each into its own container, and appended in order)
  const int order = pack_pair_order(window_modes);
  P.pair_order = order;
  const std::vector<std::vector<int32_t>> tile_rows = split_clusters(clusters, nt_max);
  std::vector<PackedHost> parts(tile_rows.size());
  parallel_chunks(tile_rows.size(), tile_rows.size(), [&](size_t b, size_t e, size_t) {
    for (size_t t = b; t < e; ++t) {
      const std::vector<int32_t>& trows = tile_rows[t];
      PackedHost& Q = parts[t];
      std::vector<uint64_t> U;
      for (int32_t r : trows) U = U.empty() ? rows[r].keys : union_of(U, rows[r].keys);
      std::vector<int32_t> where;
      layout_tile(trows, U.data(), U.size(), nin, window_modes, order, Q, where);
      const TileDesc& T = Q.tiles[0];
      Q.vals.assign((size_t)Q.vals_count, 0.0);
      double* V = Q.vals.data();
      for (int n = 0; n < T.nrows; ++n) {
        const Row& rw = rows[trows[n]];
        size_t pos = 0;
        for (size_t e2 = 0; e2 < rw.keys.size(); ++e2) {
          while (U[pos] != rw.keys[e2]) ++pos;  // both sorted; rw.keys is a subset of U
          const int at = where[pos];
          const int ks = at / 4, kk = at % 4, nt = n / 8, nn = n % 8;
          V[((size_t)ks * T.nt + nt) * 32 + nn * 4 + kk] = rw.vals[e2];  // lane = (n%8)*4 + (k%4)
        }
      }
    }
  });
  tm.lap("emit tiles");
  P = finalize_pack(std::move(P), parts, empties, nt_max);
  if (getenv("CA_TIMING") && P.pair_products > 0)
    fprintf(stderr, "[ca timing] pack_terms: %.1f %% of the pair products in (4 first factors) x (second factor) blocks\n",
            100.0 * (double)P.blocked_pairs / (double)P.pair_products);
  tm.lap("schedule");
  {  // the row lists and the term list are gigabytes at m ~ 3000: release them here, inside the timed phases
    std::vector<Row>().swap(rows);
    std::vector<Term>().swap(terms);
  }
  tm.lap("release");
  return P;
}

int decide_nt_max(const std::vector<std::vector<int32_t>>& clusters, int32_t m, int nt_max) {
  if (nt_max >= 3) {
    // Three n-tiles only where they pay: measured at m = 2962 (86 % of the rows in clusters of 17-24 rows) the windowed
    // kernel gains 13 % (f) and 28 % (Jacobian action), at m = 999 (50 %) it loses 14 % to the single CTA per SM.
    const char* force = getenv("CA_NT3");
    int64_t wide_rows = 0;
    for (auto& cl : clusters)
      if (cl.size() > 16 && cl.size() <= 24) wide_rows += (int64_t)cl.size();
    if (!(force && force[0] != '0') && (double)wide_rows < 0.7 * (double)m) nt_max = 2;
  }
  return nt_max;
}

std::vector<std::vector<int32_t>> split_clusters(const std::vector<std::vector<int32_t>>& clusters, int nt_max) {
  std::vector<std::vector<int32_t>> tile_rows;
  for (auto& cl : clusters) {
    const int n = (int)cl.size();
    const int slots = (n + 7) / 8;                     // n-tiles needed
    const int ntiles = (slots + nt_max - 1) / nt_max;  // tiles needed
    int done = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int remaining_tiles = ntiles - t;
      const int take = (n - done + remaining_tiles - 1) / remaining_tiles;  // even split, <= 8*nt_max
      tile_rows.emplace_back(cl.begin() + done, cl.begin() + done + take);
      done += take;
    }
  }
  return tile_rows;
}

// Order of the pairs inside a tile (see layout_tile):
//   kOrderAMajor    runs of equal first factor, four consecutive pairs per k-step (windowed packs: blocks make windows
//                   of fewer k-steps, and the per-window costs outweigh the saved shared-memory wavefronts)
//   kOrderBlocked   (4 first factors) x (one second factor) k-steps where the second-factor lists coincide
//   kOrderFactored  a-major with every run padded to whole k-steps, so that a k-step has ONE first factor: the
//                   quadratic evaluation can then run as  out_i += x_a * (sum_b V_iab x_b)  without pair products
//                   (bilinear_warptile_kernel<.., FACT>).  Measured at 307 modes (round 2): the runs are only ~4 k-steps
//                   long, so the per-run DFMAs cost what the DMULs did, and the padding adds 8.7 % DMMAs: 80.7 ms
//                   against 67.3 ms for the blocked order.  Kept as an option, not the default.
// Default: blocked for whole-tile packs, a-major for windowed ones.  CA_PACK_AMAJOR / CA_PACK_BLOCKED /
// CA_PACK_FACTORED force one.
int pack_pair_order(int window_modes) {
  if (getenv("CA_PACK_BLOCKED")) return kOrderBlocked;
  if (getenv("CA_PACK_AMAJOR")) return kOrderAMajor;
  if (getenv("CA_PACK_FACTORED") && getenv("CA_PACK_FACTORED")[0] != '0') return kOrderFactored;
  return window_modes <= 0 ? kOrderBlocked : kOrderAMajor;
}

void layout_tile(const std::vector<int32_t>& trows, const uint64_t* U_ptr, size_t nU, int32_t nin, int window_modes,
                 int order, PackedHost& P, std::vector<int32_t>& where_out) {
  const bool blocked = order == kOrderBlocked, factored = order == kOrderFactored && window_modes <= 0;
  struct USpan {  // read-only view with the vector interface the layout code uses
    const uint64_t* p;
    size_t n;
    size_t size() const { return n; }
    const uint64_t& operator[](size_t q) const { return p[q]; }
  };
  const USpan U{U_ptr, nU};
  {
    TileDesc T{};
    T.nrows = (int32_t)trows.size();
    T.nt = (T.nrows + 7) / 8;
    T.npairs = (int32_t)U.size();
    T.ks_begin = 0;
    T.vals_off = 0;
    T.rows_off = 0;
    for (int n = 0; n < 8 * T.nt; ++n) P.rows.push_back(n < T.nrows ? trows[n] : -1);
    const uint32_t ONE = (uint32_t)(nin - 1);
    const int cap = window_modes > 0 ? window_modes - 1 : 1 << 30;  // leave room for the padding pseudo mode
    const size_t max_pairs = window_modes > 0 ? (size_t)4 * kWindowMaxKs : ~(size_t)0;

    // ---- emission order of the pairs, as SEGMENTS (units that a window never splits).
    // Blocked segments first: first factors whose second-factor lists coincide (the modes of one wavenumber group and
    // parity all pair with the same modes) are taken four at a time, and a k-step holds (a0,b), (a1,b), (a2,b), (a3,b)
    // for one b: the four a stay in registers over the whole run of b (kSameA), and the four lanes of a state read the
    // SAME second factor -- one shared-memory wavefront instead of two per load.  A segment is (quads of a) x (chunk
    // of b), sized so that it fits a window: q quads and nb second factors touch only 4q + nb modes for q * nb k-steps.
    // What does not fit the pattern (first factors left over, triangular parts) follows a-major, four consecutive
    // pairs per k-step as before.
    struct Seg { std::vector<size_t> idx; };
    std::vector<Seg> segs;
    std::vector<char> taken(U.size(), 0);
    if (blocked && U.size() >= 16 && cap >= 11) {
      // runs of equal first factor
      struct Run { uint32_t a; size_t b, e; uint64_t sig; };
      std::vector<Run> runs;
      for (size_t r = 0; r < U.size();) {
        size_t e = r;
        const uint32_t a = (uint32_t)(U[r] / nin);
        uint64_t sig = 0x9e3779b97f4a7c15ull;
        while (e < U.size() && (uint32_t)(U[e] / nin) == a) sig = mix64(sig ^ (U[e++] % nin));
        runs.push_back({a, r, e, sig});
        r = e;
      }
      std::vector<size_t> order(runs.size());
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) {
        if (runs[x].e - runs[x].b != runs[y].e - runs[y].b) return runs[x].e - runs[x].b > runs[y].e - runs[y].b;
        return runs[x].sig < runs[y].sig;
      });
      auto same_list = [&](const Run& x, const Run& y) {
        if (x.e - x.b != y.e - y.b || x.sig != y.sig) return false;
        for (size_t q = 0; q < x.e - x.b; ++q)
          if (U[x.b + q] % nin != U[y.b + q] % nin) return false;
        return true;
      };
      for (size_t c0 = 0; c0 < order.size();) {
        size_t c1 = c0 + 1;
        while (c1 < order.size() && same_list(runs[order[c0]], runs[order[c1]])) ++c1;
        const size_t nq = (c1 - c0) / 4, nb = runs[order[c0]].e - runs[order[c0]].b;
        if (nq >= 1 && nb >= 2) {
          std::vector<size_t> cls(order.begin() + (ptrdiff_t)c0, order.begin() + (ptrdiff_t)(c0 + 4 * nq));
          std::sort(cls.begin(), cls.end());  // ascending first factor
          // window shape: nbc second factors x qw quads with 4 qw + nbc <= cap and qw * nbc <= kWindowMaxKs
          size_t nbc = nb, qw = nq;
          if (window_modes > 0) {  // near-equal chunks of <= min(cap / 2, kWindowMaxKs / 4) second factors
            const size_t lim = std::max<size_t>(1, std::min<size_t>((size_t)cap / 2, (size_t)kWindowMaxKs / 4));
            const size_t nchunks = (nb + lim - 1) / lim;
            nbc = (nb + nchunks - 1) / nchunks;
            qw = std::max<size_t>(1, std::min<size_t>({nq, ((size_t)cap - nbc) / 4, (size_t)kWindowMaxKs / nbc}));
          }
          for (size_t b0 = 0; b0 < nb; b0 += nbc)
            for (size_t q0 = 0; q0 < nq; q0 += qw) {
              Seg sg;
              for (size_t q = q0; q < std::min(nq, q0 + qw); ++q)
                for (size_t bb = b0; bb < std::min(nb, b0 + nbc); ++bb)
                  for (int c = 0; c < 4; ++c) {
                    const size_t at = runs[cls[4 * q + c]].b + bb;
                    sg.idx.push_back(at);
                    taken[at] = 1;
                  }
              segs.push_back(std::move(sg));
            }
        }
        c0 = c1;
      }
    }
    {  // a-major remainder: runs of equal first factor, split where a run alone exceeds a window
      const size_t run_cap = window_modes > 0 ? (size_t)std::max(cap - 1, 1) : ~(size_t)0;
      for (size_t r = 0; r < U.size();) {
        size_t e = r;
        const uint32_t a = (uint32_t)(U[r] / nin);
        Seg sg;
        while (e < U.size() && (uint32_t)(U[e] / nin) == a) {
          if (!taken[e]) {
            if (sg.idx.size() == run_cap) {
              segs.push_back(std::move(sg));
              sg = Seg();
            }
            sg.idx.push_back(e);
          }
          ++e;
        }
        if (!sg.idx.empty()) segs.push_back(std::move(sg));
        r = e;
      }
    }

    for (char t : taken) P.blocked_pairs += t;
    // position of every pair of U in the emitted (padded) pair sequence
    std::vector<int32_t> where(U.size());
    auto mark_same_a = [&](size_t first_word, int nks) {
      for (int ks = 1; ks < nks; ++ks) {
        uint32_t* cur = &P.pairs[first_word + (size_t)ks * 4];
        const uint32_t* prev = cur - 4;
        bool same = true;
        for (int c = 0; c < 4; ++c) same = same && ((cur[c] & 0x7fffu) == (prev[c] & 0x7fffu));
        if (same)
          for (int c = 0; c < 4; ++c) cur[c] |= 0x8000u;
      }
    };
    if (window_modes <= 0) {
      int k = 0;
      for (const Seg& sg : segs) {
        for (size_t at : sg.idx) {
          where[at] = k++;
          P.pairs.push_back((uint32_t)(U[at] / nin) | ((uint32_t)(U[at] % nin) << 16));
        }
        if (factored && !sg.idx.empty())  // whole k-steps per first factor; the padding pairs (a, ONE) carry zero values
          for (const uint32_t a = (uint32_t)(U[sg.idx[0]] / nin); k % 4 != 0; ++k) P.pairs.push_back(a | (ONE << 16));
      }
      T.nks = (k + 3) / 4;
      for (; k < 4 * T.nks; ++k) P.pairs.push_back(ONE | (ONE << 16));
      mark_same_a(0, T.nks);
    } else {
      // ---- windows: segments are appended while the window's mode set and k-step count stay within the caps
      std::vector<uint32_t> wmodes;  // sorted global modes of the open window
      std::vector<size_t> wpairs;    // indices into U
      int emitted = 0;               // pairs emitted so far in this tile (incl. padding)
      auto close_window = [&]() {
        if (wpairs.empty()) return;
        const bool pad = wpairs.size() % 4 != 0;
        if (pad && !std::binary_search(wmodes.begin(), wmodes.end(), ONE)) {
          wmodes.push_back(ONE);
          std::sort(wmodes.begin(), wmodes.end());
        }
        WindowDesc W{};
        W.ks_begin = emitted / 4;
        W.nks = (int32_t)((wpairs.size() + 3) / 4);
        W.modes_off = (int32_t)P.modes.size();
        W.nmodes = (int32_t)wmodes.size();
        W.tile = 0;
        W.last = 0;
        for (uint32_t g : wmodes) P.modes.push_back((int32_t)g);
        auto local = [&](uint32_t g) { return (uint32_t)(std::lower_bound(wmodes.begin(), wmodes.end(), g) - wmodes.begin()); };
        const size_t first_word = P.pairs.size();
        for (size_t q = 0; q < (size_t)W.nks * 4; ++q) {
          if (q < wpairs.size()) {
            const uint64_t key = U[wpairs[q]];
            where[wpairs[q]] = emitted + (int32_t)q;
            P.pairs.push_back(local((uint32_t)(key / nin)) | (local((uint32_t)(key % nin)) << 16));
          } else {
            P.pairs.push_back(local(ONE) | (local(ONE) << 16));
          }
        }
        mark_same_a(first_word, W.nks);
        emitted += W.nks * 4;
        P.max_window_modes = std::max(P.max_window_modes, W.nmodes);
        P.windows.push_back(W);
        wmodes.clear();
        wpairs.clear();
      };
      for (const Seg& sg : segs) {
        std::vector<uint32_t> add;
        for (size_t at : sg.idx) {
          add.push_back((uint32_t)(U[at] / nin));
          add.push_back((uint32_t)(U[at] % nin));
        }
        std::sort(add.begin(), add.end());
        add.erase(std::unique(add.begin(), add.end()), add.end());
        std::vector<uint32_t> u;
        std::set_union(wmodes.begin(), wmodes.end(), add.begin(), add.end(), std::back_inserter(u));
        // (blocked segments are whole k-steps and all come before the a-major remainder, so they stay aligned)
        if (!wpairs.empty() && ((int)u.size() > cap || wpairs.size() + sg.idx.size() > max_pairs)) {
          close_window();
          u = add;
        }
        wmodes = std::move(u);
        wpairs.insert(wpairs.end(), sg.idx.begin(), sg.idx.end());
      }
      close_window();
      if (emitted > 0) P.windows.back().last = 1;
      T.nks = emitted / 4;
    }
    P.vals_count = (int64_t)T.nks * T.nt * 32;
    P.padded_fma += (int64_t)T.nks * 4 * T.nt * 8;
    P.pair_products += (int64_t)T.nks * 4;
    P.tiles.push_back(T);
    where_out.swap(where);
  }
}

PackedHost finalize_pack(PackedHost P, std::vector<PackedHost>& parts, const std::vector<int32_t>& empties, int nt_max) {
  const int window_modes = P.window_modes;
  const bool have_vals = !parts.empty() && (int64_t)parts[0].vals.size() == parts[0].vals_count;
  {
    int64_t vals_total = 0;
    for (PackedHost& Q : parts) {
      const int32_t ks_base = (int32_t)(P.pairs.size() / 4), rows_base = (int32_t)P.rows.size();
      const int32_t modes_base = (int32_t)P.modes.size(), tile_index = (int32_t)P.tiles.size();
      const int64_t vals_base = vals_total;
      TileDesc T = Q.tiles[0];
      T.ks_begin += ks_base;
      T.vals_off += vals_base;
      T.rows_off += rows_base;
      P.tiles.push_back(T);
      for (WindowDesc W : Q.windows) {
        W.ks_begin += ks_base;
        W.modes_off += modes_base;
        W.tile = tile_index;
        P.windows.push_back(W);
      }
      P.pairs.insert(P.pairs.end(), Q.pairs.begin(), Q.pairs.end());
      if (have_vals) P.vals.insert(P.vals.end(), Q.vals.begin(), Q.vals.end());
      vals_total += Q.vals_count;
      P.rows.insert(P.rows.end(), Q.rows.begin(), Q.rows.end());
      P.modes.insert(P.modes.end(), Q.modes.begin(), Q.modes.end());
      P.max_window_modes = std::max(P.max_window_modes, Q.max_window_modes);
      P.padded_fma += Q.padded_fma;
      P.pair_products += Q.pair_products;
      P.blocked_pairs += Q.blocked_pairs;
      Q = PackedHost();  // free the tile's copy
    }
    P.vals_count = vals_total;
  }
  // rows without any term still have to be written (as zeros)
  for (size_t e = 0; e < empties.size(); e += 8 * (size_t)nt_max) {
    size_t n = std::min(empties.size() - e, (size_t)8 * nt_max);
    TileDesc T{};
    T.nrows = (int32_t)n;
    T.nt = (T.nrows + 7) / 8;
    T.nks = 0;
    T.npairs = 0;
    T.ks_begin = (int32_t)(P.pairs.size() / 4);
    T.vals_off = P.vals_count;
    T.rows_off = (int32_t)P.rows.size();
    for (int k = 0; k < 8 * T.nt; ++k) P.rows.push_back(k < T.nrows ? empties[e + k] : -1);
    P.tiles.push_back(T);
  }
  // wide tiles first (the kernel runs one code path per nt), long before short inside each width
  std::vector<int32_t> perm(P.tiles.size());
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int32_t x, int32_t y) {
    if (P.tiles[x].nt != P.tiles[y].nt) return P.tiles[x].nt > P.tiles[y].nt;
    return P.tiles[x].nks > P.tiles[y].nks;
  });
  std::vector<TileDesc> sorted;
  sorted.reserve(P.tiles.size());
  for (int32_t o : perm) sorted.push_back(P.tiles[o]);
  if (window_modes > 0) {
    std::vector<std::vector<WindowDesc>> by_tile(P.tiles.size());
    for (const WindowDesc& W : P.windows) by_tile[W.tile].push_back(W);
    P.windows.clear();
    for (size_t t = 0; t < perm.size(); ++t) {
      std::vector<WindowDesc>& ws = by_tile[perm[t]];
      if (ws.empty()) {  // a tile without terms still has to write its zeros
        WindowDesc W{};
        W.ks_begin = sorted[t].ks_begin;
        W.last = 1;
        ws.push_back(W);
      }
      for (WindowDesc W : ws) {
        W.tile = (int32_t)t;
        W.rows_off = sorted[t].rows_off;
        W.nt = sorted[t].nt;
        W.vals_off = sorted[t].vals_off + (int64_t)(W.ks_begin - sorted[t].ks_begin) * sorted[t].nt * 32;
        P.windows.push_back(W);
      }
    }
  }
  P.tiles = std::move(sorted);
  if (window_modes > 0 && getenv("CA_TIMING")) {
    int64_t ks = 0, md = 0;
    int64_t hist[5] = {0, 0, 0, 0, 0}, hks[5] = {0, 0, 0, 0, 0};  // windows / k-steps by window length: < 64, < 128, < 256, < 384, more
    for (const WindowDesc& W : P.windows) {
      ks += W.nks;
      md += W.nmodes;
      const int b = W.nks < 64 ? 0 : W.nks < 128 ? 1 : W.nks < 256 ? 2 : W.nks < 384 ? 3 : 4;
      hist[b]++;
      hks[b] += W.nks;
    }
    const double n = std::max<double>(1.0, (double)P.windows.size());
    fprintf(stderr, "[ca timing] pack_terms: %zu tiles, %zu windows, %.1f k-steps and %.1f modes per window; windows (k-steps) by length "
            "< 64: %lld (%lld), < 128: %lld (%lld), < 256: %lld (%lld), < 384: %lld (%lld), more: %lld (%lld)\n",
            P.tiles.size(), P.windows.size(), ks / n, md / n, (long long)hist[0], (long long)hks[0], (long long)hist[1], (long long)hks[1],
            (long long)hist[2], (long long)hks[2], (long long)hist[3], (long long)hks[3], (long long)hist[4], (long long)hks[4]);
  }
  {  // LPT schedule of whole tiles onto 8 warps (weight = DMMAs per state tile)
    std::vector<int32_t> idx(P.tiles.size());
    std::iota(idx.begin(), idx.end(), 0);
    // cycles of the FP64 pipe per k-step: 4 nt DMMAs x 16 plus the four pair products (tools/dmma_dmul_mix.cu)
    // (a one-n-tile k-step, 4 DMMA + 4 DMUL, takes 0.78 of the time of a two-n-tile one in the kernel -- its phases keep
    // the pipe only ~70 % busy; measured with weights 82 / 100 / 112 / 125: 13.38 / 13.30 / 13.50 / 13.73 ms per 2*10^5
    // states at 307 modes -- the schedule's granularity matters more than the exact weight)
    auto weight = [&](int32_t t) { return (int64_t)P.tiles[t].nks * (P.tiles[t].nt == 2 ? 144 : 100) + 40; };
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) { return weight(x) > weight(y); });
    int64_t load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<int32_t> bins[8];
    for (int32_t t : idx) {
      int w = 0;
      for (int q = 1; q < 8; ++q)
        if (load[q] < load[w]) w = q;
      load[w] += weight(t);
      bins[w].push_back(t);
    }
    int64_t total = 0, worst = 0;
    for (int w = 0; w < 8; ++w) {
      total += load[w];
      worst = std::max(worst, load[w]);
      std::stable_sort(bins[w].begin(), bins[w].end(), [&](int32_t x, int32_t y) { return P.tiles[x].nt > P.tiles[y].nt; });
      P.sched_off[w] = (int32_t)P.sched.size();
      P.sched_nt2[w] = 0;
      for (int32_t t : bins[w]) {
        P.sched.push_back(t);
        P.sched_nt2[w] += P.tiles[t].nt >= 2;
      }
    }
    P.sched_off[8] = (int32_t)P.sched.size();
    P.sched_imbalance = total > 0 ? (double)worst * 8.0 / (double)total : 1e9;
    if (window_modes <= 0) {  // physical re-layout in schedule order: one contiguous stream per warp and nt class
      std::vector<uint32_t> np;
      std::vector<double> nv;
      np.reserve(P.pairs.size());
      if (have_vals) nv.reserve(P.vals.size());
      int64_t vals_at = 0;
      for (int w = 0; w < 8; ++w) {
        P.sched_ks2[w] = P.sched_ks1[w] = 0;
        for (int q = P.sched_off[w]; q < P.sched_off[w + 1]; ++q) {
          TileDesc& T = P.tiles[P.sched[q]];
          const size_t p0 = (size_t)T.ks_begin * 4, v0 = (size_t)T.vals_off;
          // every tile starts on a batch boundary of the kernel's pair ring (kStreamAlign k-steps): the k-step loop then
          // runs in whole batches with one ring refill per batch and no branch inside
          np.resize((np.size() + 4 * kStreamAlign - 1) / (4 * kStreamAlign) * (4 * kStreamAlign), 0u);
          T.ks_begin = (int32_t)(np.size() / 4);
          T.vals_off = vals_at;
          vals_at += (int64_t)T.nks * T.nt * 32;
          np.insert(np.end(), P.pairs.begin() + p0, P.pairs.begin() + p0 + (size_t)T.nks * 4);
          if (have_vals) nv.insert(nv.end(), P.vals.begin() + v0, P.vals.begin() + v0 + (size_t)T.nks * T.nt * 32);
          (T.nt == 2 ? P.sched_ks2[w] : P.sched_ks1[w]) += (T.nks + kStreamAlign - 1) / kStreamAlign * kStreamAlign;
        }
      }
      np.resize((np.size() + 4 * kStreamAlign - 1) / (4 * kStreamAlign) * (4 * kStreamAlign), 0u);
      P.pairs.swap(np);
      if (have_vals) P.vals.swap(nv);
      P.stream_layout = true;
    }
  }
  return P;
}

}  // namespace ca
