// rsd_long.inl — host side of the long-pair path, second generation (included by rsd_api.cu).
//
//   rsd_long_pairs   a batch of long pairs in as few cooperative launches as memory allows: pairs are spread over
//                    rings of CTAs (k_long2.cuh), traceback and script emission run one CTA per pair.
//   overflow path    a pair whose direction matrix (2 bit per cell) plus boundary columns exceed the memory budget,
//                    or whose column panels exceed the co-resident CTAs, is cut into row blocks x panel ranges:
//                    pass 1 runs the blocks top to bottom without directions and keeps the key row at every block
//                    boundary (4 bytes per column), pass 2 recomputes the blocks bottom to top with directions and
//                    traces back through each.  Linear space in the Hirschberg sense (O(n) per checkpoint row,
//                    ~2x the cell updates); forward checkpoints instead of Hirschberg's backward half because the
//                    reference's canonical tie-break (first of insert, delete, update among the (cost, steps)-minimal
//                    predecessors, SED:244-265) is defined on forward keys — the recomputed blocks reproduce exactly the
//                    direction words of the one-launch path.
// Pairs that need the exact-double or fp64 kernels (non-dyadic costs, very large costs) go through rsd_long_pair.

namespace {

struct LongPairPlan {
    int64_t m = 0, n = 0;
    int S = 0, n_panels = 0;
    int64_t n_pad = 0;
    bool eligible = false, trivial = false;
    int64_t hb = 0;                // rows per block (multiple of 32, or m when one block)
    int nb = 1, nr = 1;            // row blocks, panel ranges
    size_t need = 0;               // device bytes
    // device pointers (inside the pool)
    uint8_t *da = nullptr, *db = nullptr, *tmp = nullptr, *op = nullptr;
    int32_t *oi = nullptr, *oj = nullptr;
    uint32_t *dirs = nullptr, *ckpt = nullptr;
    unsigned long long *bound = nullptr;
    long long *keyacc = nullptr;   // [0] real, [1] scratch of the recomputation pass
    double *dist = nullptr;
    int *state = nullptr; int32_t *n_ops = nullptr;
};

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// boundary columns live at the head of the pool, a region that only ever holds tagged boundary words, so a stale word can
// never pass for a published one (see LongJob2::gen and rsd_ctx::long_bound_hw)
inline size_t long_bound_bytes(const LongPairPlan &P) { return al256((size_t)P.n_panels * (size_t)P.hb * 8); }

// carve the pair's other buffers out of [base, ...); returns the bytes used (base == nullptr: size only)
size_t long_layout(LongPairPlan &P, unsigned char *base, bool want_script, bool want_ij) {
    (void)want_ij;
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char *p = base ? base + off : nullptr; off += al256(bytes); return p; };
    if (want_script) {
        P.dirs = (uint32_t *)take((size_t)long2_dir_groups(P.hb) * (size_t)P.n_pad * 4 + 64);
        P.tmp = (uint8_t *)take((size_t)(P.m + P.n) + 64);
    }
    if (P.nb > 1) P.ckpt = (uint32_t *)take((size_t)(want_script ? P.nb - 1 : 2) * (size_t)P.n_pad * 4);
    return off;
}
// The pairs of a batch share four contiguous regions so that inputs and results cross PCIe in a handful of copies through
// a pinned staging buffer instead of five small pageable copies per pair: sequences (a then b of every pair), 64 bytes
// of scalars per pair (exact key accumulators, distance, traceback state, op count), and the packed scripts (op, oi, oj).
inline size_t long_seq_bytes(const LongPairPlan &P) { return al256((size_t)P.m + 64) + al256((size_t)P.n + 64); }
inline size_t long_ops_slot(const LongPairPlan &P) { return al256((size_t)(P.m + P.n) + 64); }                 // entries, not bytes
inline size_t long_shared_bytes(const LongPairPlan &P, bool want_script, bool want_ij) {
    return long_seq_bytes(P) + 64 + (want_script ? long_ops_slot(P) * (want_ij ? 9 : 1) : 0);
}

}  // namespace

extern "C" int rsd_long_pairs(rsd_ctx *c, int n_pairs, const uint8_t *const *a, const int64_t *m, const uint8_t *const *b, const int64_t *n,
                              int force_mode, int want_script, const int64_t *max_ops,
                              uint8_t *const *op, int32_t *const *oi, int32_t *const *oj, int64_t *n_ops, double *dist, int *mode_out) {
    if (!c) return rsd_fail(RSD_EINVAL, "ctx is NULL");
    if (n_pairs < 0 || (n_pairs > 0 && (!a || !m || !b || !n || !dist))) return rsd_fail(RSD_EINVAL, "rsd_long_pairs: bad arguments");
    if (want_script && n_pairs > 0 && (!op || !n_ops || !max_ops)) return rsd_fail(RSD_EINVAL, "rsd_long_pairs: script buffers missing");
    RSD_OK_OR_RETURN(c->ensure_device());
    if (n_pairs == 0) return RSD_OK;
    const bool want_ij = want_script && (oi || oj);
    const auto t_entry = std::chrono::steady_clock::now();
    auto since_ms = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    int C = 4;
    std::vector<LongPairPlan> P((size_t)n_pairs);
    std::vector<int> fallback;                       // pairs for rsd_long_pair (fp64 / exact-double keys)
    ModeInfo mi_up{}; bool have_mi = false;
    long long maxc_all = 0;                          // largest |scaled cost| among the symbols of the call
    // one set of integer cost tables for the whole call: classified on the union of the symbols of every pair
    uint32_t symmask = 0;
    for (int p = 0; p < n_pairs; ++p) {
        LongPairPlan &Q = P[p];
        Q.m = m[p]; Q.n = n[p];
        if (Q.m < 0 || Q.n < 0 || (Q.m > 0 && !a[p]) || (Q.n > 0 && !b[p])) return rsd_fail(RSD_EINVAL, "rsd_long_pairs: bad pair %d", p);
        if (Q.m > 0x3fffffff || Q.n > 0x3fffffff) return rsd_fail(RSD_ERANGE, "rsd_long_pairs: sequence too long");
        if (want_script && (!op[p] || max_ops[p] < Q.m + Q.n)) return rsd_fail(RSD_EINVAL, "rsd_long_pairs: pair %d: script buffer missing or max_ops < m+n", p);
        if (Q.m == 0 || Q.n == 0) { Q.trivial = true; continue; }
        uint32_t sm = 0;
        for (int side = 0; side < 2; ++side) {                         // four independent accumulators: ~0.4 ns per symbol
            const uint8_t *x = side ? b[p] : a[p]; const int64_t len = side ? Q.n : Q.m;
            uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0; unsigned hi = 0;
            int64_t i = 0;
            for (; i + 4 <= len; i += 4) {
                hi |= x[i] | x[i + 1] | x[i + 2] | x[i + 3];
                m0 |= 1u << (x[i] & 31); m1 |= 1u << (x[i + 1] & 31); m2 |= 1u << (x[i + 2] & 31); m3 |= 1u << (x[i + 3] & 31);
            }
            for (; i < len; ++i) { hi |= x[i]; m0 |= 1u << (x[i] & 31); }
            if (hi > 15) return rsd_fail(RSD_EINVAL, "rsd_long_pairs: code > 15");
            sm |= m0 | m1 | m2 | m3;
        }
        symmask |= sm;
    }
    // Rings and panel width, from measurements on 50 kb pairs (tools/dbg_long_batch.py, profiles/r02_long_batch_sweep.log):
    // a round of r pairs side by side (one ring each, 16 columns per lane) takes 7.6 / 11.3 / 11.0 / 16.5 ms for
    // r = 4 / 8 / 12 / 16, i.e. 1.9 / 1.4 / 0.92 / 1.03 ms per pair: about 1200 warps in flight (2 per SM scheduler) is
    // the sweet spot, more only adds contention.  So: r* = 1200 / (panels of a pair at 16 columns per lane) rings when
    // the batch has clearly more pairs than that (more than r* + r*/3), else one ring per pair (a second round would cost
    // more than the contention); up to 128 rings, so batches of 5 - 20 kb pairs reach the same occupancy; 16 columns per lane once ~700 warps are in flight at 8, else 8 (a lone warp's step
    // hardly grows from 4 to 8 columns, so 4 never wins).
    int rings_want = 1;
    {
        int nk = 0; double sum_n = 0;
        for (int p = 0; p < n_pairs; ++p) if (!P[p].trivial) { ++nk; sum_n += (double)P[p].n; }
        const double avg_n = nk ? sum_n / nk : 0.0;
        const int r_star = (int)std::min((double)RSD_LONG2_MAX_RINGS, std::max(1.0, std::floor(1200.0 / std::max(1.0, std::ceil(avg_n / 512.0)) + 0.5)));
        rings_want = nk <= std::min(RSD_LONG2_MAX_RINGS, r_star + r_star / 3) ? std::max(nk, 1) : r_star;
        if (const char *e = getenv("RSD_LONG_RINGS")) rings_want = std::min(std::max(atoi(e), 1), RSD_LONG2_MAX_RINGS);
        C = (double)std::min(nk, rings_want) * (avg_n / 256.0) >= 700.0 ? 16 : 8;
        if (const char *e = getenv("RSD_LONG_C")) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16) C = v; }
    }
    if (symmask) {
        ModeInfo mi;
        RSD_OK_OR_RETURN(c->classify(symmask, 1, 1, 4, RSD_MODE_F64, mi));       // the tables; the key width is checked per pair below
        const bool int_ok = mi.dyadic && force_mode != RSD_MODE_F64 && !getenv("RSD_LONG_WIDE");
        long long maxc = std::max<long long>(mi.ic.ins, mi.ic.del);
        for (int x = 0; x < 16; ++x) for (int y = 0; y < 16; ++y)
            if ((symmask >> x & 1) && (symmask >> y & 1)) maxc = std::max<long long>(maxc, std::llabs((long long)mi.ic.w[x][y]));
        maxc_all = maxc;
        for (int p = 0; p < n_pairs; ++p) {
            LongPairPlan &Q = P[p];
            if (Q.trivial) continue;
            Q.S = ceil_log2_i64(Q.m + Q.n + 66);
            // test knob: a wider steps field makes the 32-bit modular keys wrap on small matrices (see rsd_long_pair)
            if (const char *e = getenv("RSD_LONG_S")) Q.S = std::min(std::max(Q.S, atoi(e)), 30);
            bool ok32 = int_ok && Q.S <= 24 && ((64 * maxc + 128) << Q.S) < (1ll << 30);     // 32 rows of drift stay below 2^30
            const double bound = ((double)Q.m * mi.ic.del + (double)(Q.n + 512) * mi.ic.ins + 4.0 * ((double)mi.ic.ins + mi.ic.del)) * std::ldexp(1.0, Q.S);
            if (bound > 4.0e18) ok32 = false;                                  // the exact key of (m, n) must fit an int64
            if (!ok32) { fallback.push_back(p); continue; }
            Q.eligible = true; have_mi = true;
            if (mode_out) mode_out[p] = RSD_MODE_I32;
            Q.n_panels = (int)((Q.n + 32 * C - 1) / (32 * C));
            Q.n_pad = (int64_t)Q.n_panels * 32 * C;
        }
        mi_up = mi;
    }
    c->timed = false; c->last_ms_override = 0.0;
    cudaStream_t st = c->stream;
    // ---- trivial pairs: border row / column only (SED:146-182) ----
    for (int p = 0; p < n_pairs; ++p) if (P[p].trivial) {
        const int64_t mm = P[p].m, nn = P[p].n;
        dist[p] = mm == 0 ? (double)nn * c->ins : (double)mm * c->del;
        if (mode_out) mode_out[p] = RSD_MODE_I32;
        if (want_script) {
            const int64_t k = mm + nn;
            for (int64_t x = 0; x < k; ++x) { op[p][x] = mm == 0 ? 0 : 1; if (oi && oi[p]) oi[p][x] = mm == 0 ? 0 : (int32_t)(x + 1); if (oj && oj[p]) oj[p][x] = mm == 0 ? (int32_t)(x + 1) : 0; }
            n_ops[p] = k;
        }
    }
    if (have_mi) {
        RSD_OK_OR_RETURN(c->upload_costs(mi_up, st));
        // relative keys (plain signed compares) when every pair's keys stay within 2^30 of their lane's base: (C + 40)
        // border steps of at most maxc << S each — implied by the eligibility bound above for C <= 16, checked anyway;
        // RSD_LONG_NOREL keeps the wrapped-difference kernels under test
        bool rel = !getenv("RSD_LONG_NOREL");
        for (int p = 0; p < n_pairs; ++p) if (P[p].eligible && (((long long)(C + 40) * maxc_all) << P[p].S) >= (1ll << 30)) rel = false;
        const void *kfn_d = rel ? (C == 4 ? (const void *)k_long2<4, true, true> : C == 8 ? (const void *)k_long2<8, true, true> : (const void *)k_long2<16, true, true>)
                                : (C == 4 ? (const void *)k_long2<4, true, false> : C == 8 ? (const void *)k_long2<8, true, false> : (const void *)k_long2<16, true, false>);
        const void *kfn_n = rel ? (C == 4 ? (const void *)k_long2<4, false, true> : C == 8 ? (const void *)k_long2<8, false, true> : (const void *)k_long2<16, false, true>)
                                : (C == 4 ? (const void *)k_long2<4, false, false> : C == 8 ? (const void *)k_long2<8, false, false> : (const void *)k_long2<16, false, false>);
        int per_sm = 0;
        RSD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn_d, 32, 0));
        int64_t max_ctas = (int64_t)per_sm * c->sm_count;
        if (const char *e = getenv("RSD_LONG_MAXCTAS")) max_ctas = std::max<int64_t>(1, std::min<int64_t>(max_ctas, atoll(e)));
        // memory budget: what is free now plus what this context already holds for long pairs
        // (cudaMemGetInfo costs ~15 ms per call on this driver: asked once per context, and again after a failed allocation)
        if (c->long_budget == 0) {
            size_t free_b = 0, total_b = 0;
            RSD_CUDA(cudaMemGetInfo(&free_b, &total_b));
            c->long_budget = (size_t)((double)(free_b + c->long_pool.cap) * 0.85);
        }
        size_t budget = c->long_budget;
        if (const char *e = getenv("RSD_LONG_BUDGET_MB")) budget = (size_t)atoll(e) << 20;
        // ---- per pair: one block if it fits, else row blocks x panel ranges ----
        for (int p = 0; p < n_pairs; ++p) if (P[p].eligible) {
            LongPairPlan &Q = P[p];
            Q.hb = Q.m; Q.nb = 1; Q.nr = (int)((Q.n_panels + max_ctas - 1) / max_ctas);
            Q.need = long_layout(Q, nullptr, want_script != 0, want_ij) + long_bound_bytes(Q) + long_shared_bytes(Q, want_script != 0, want_ij);
            if (Q.need > budget) {
                LongPairPlan T = Q;
                T.hb = 32; T.nb = 2;
                const size_t fixed = long_layout(T, nullptr, want_script != 0, want_ij) + long_bound_bytes(T) + long_shared_bytes(T, want_script != 0, want_ij);       // everything but the per-row parts, at 32 rows
                const size_t per_row = (size_t)Q.n_panels * 8 + (want_script ? (size_t)Q.n_pad / 4 : 0);
                // checkpoint rows: one per block boundary (4 bytes per column)
                int64_t hb = 0;
                for (int64_t try_hb = (Q.m + 31) / 32 * 32; try_hb >= 32; try_hb = (try_hb / 2 + 31) / 32 * 32) {
                    const int64_t nb = (Q.m + try_hb - 1) / try_hb;
                    const size_t tot = fixed + per_row * (size_t)try_hb + (size_t)(want_script ? nb : 2) * (size_t)Q.n_pad * 4;
                    if (tot <= budget) { hb = try_hb; break; }
                    if (try_hb == 32) break;
                }
                if (hb == 0) return rsd_fail(RSD_ENOMEM, "rsd_long_pairs: pair %d (%lld x %lld) does not fit the memory budget of %zu MB even in 32-row blocks",
                                             p, (long long)Q.m, (long long)Q.n, budget >> 20);
                // grow the block while it still fits (the halving search above stops at the first fit)
                while (true) {
                    const int64_t h2 = hb + std::max<int64_t>(32, hb / 8 / 32 * 32);
                    if (h2 >= Q.m) break;
                    const int64_t nb2 = (Q.m + h2 - 1) / h2;
                    if (fixed + per_row * (size_t)h2 + (size_t)(want_script ? nb2 : 2) * (size_t)Q.n_pad * 4 > budget) break;
                    hb = h2;
                }
                Q.hb = hb; Q.nb = (int)((Q.m + hb - 1) / hb);
                Q.need = long_layout(Q, nullptr, want_script != 0, want_ij) + long_bound_bytes(Q) + long_shared_bytes(Q, want_script != 0, want_ij);
            }
        }
        // ---- batches: consecutive eligible pairs that fit the budget together; a blocked pair runs alone ----
        std::vector<int> order;
        for (int p = 0; p < n_pairs; ++p) if (P[p].eligible) order.push_back(p);
        const bool ltrace = getenv("RSD_TRACE") != nullptr;
        size_t at = 0;
        bool first_batch = true;
        float fwd_ms_total = 0.f;
        while (at < order.size()) {
            std::vector<int> batch;
            size_t need = 0;
            const bool solo = P[order[at]].nb > 1 || P[order[at]].nr > 1;
            if (solo) { batch.push_back(order[at]); need = P[order[at]].need; ++at; }
            else while (at < order.size() && P[order[at]].nb == 1 && P[order[at]].nr == 1 && (batch.empty() || need + P[order[at]].need <= budget)) {
                need += P[order[at]].need; batch.push_back(order[at]); ++at;
            }
            size_t bound_need = 0;
            for (int p : batch) bound_need += long_bound_bytes(P[p]);
            // Boundary words are tagged with a generation number instead of being reset before every launch.  That is only
            // safe where nothing but boundary words has ever been stored: they occupy the head of the pool, [0, long_bound_hw),
            // a region that only grows; the part it grows into is zeroed once (tag 0 is never used), and a new pool starts over.
            {
                const void *before = c->long_pool.p; const size_t before_cap = c->long_pool.cap;
                const size_t head = std::max(c->long_bound_hw, bound_need);
                if (int rc = c->long_pool.ensure(head + (need - bound_need) + 4096)) { c->long_budget = 0; c->long_bound_hw = 0; return rc; }   // other allocations moved in: measure again next time
                if (c->long_pool.p != before || c->long_pool.cap != before_cap) c->long_bound_hw = 0;
                if (bound_need > c->long_bound_hw) {
                    RSD_CUDA(cudaMemsetAsync((unsigned char *)c->long_pool.p + c->long_bound_hw, 0, bound_need - c->long_bound_hw, st));
                    c->long_bound_hw = bound_need;
                }
            }
            unsigned char *bbase = (unsigned char *)c->long_pool.p, *base = bbase + c->long_bound_hw;
            size_t off = 0, boff = 0;
            // shared regions first: [sequences][scalars][op][oi][oj]
            size_t seq_bytes = 0, ops_entries = 0;
            for (int p : batch) { seq_bytes += long_seq_bytes(P[p]); ops_entries += long_ops_slot(P[p]); }
            const size_t small_bytes = al256(64 * batch.size());
            unsigned char *d_seq = base + off; off += seq_bytes;
            unsigned char *d_small = base + off; off += small_bytes;
            uint8_t *d_op = nullptr; int32_t *d_oi = nullptr, *d_oj = nullptr;
            if (want_script) {
                d_op = base + off; off += ops_entries;
                if (want_ij) { d_oi = (int32_t *)(base + off); off += 4 * ops_entries; d_oj = (int32_t *)(base + off); off += 4 * ops_entries; }
            }
            // pinned staging: sequences on the way in, scalars and scripts on the way out
            const size_t out_bytes = small_bytes + (want_script ? ops_entries * (want_ij ? 9 : 1) : 0);
            const size_t stage_need = std::max(seq_bytes, out_bytes) + 256;
            if (stage_need > c->long_hstage_cap) {
                if (c->long_hstage) cudaFreeHost(c->long_hstage);
                c->long_hstage = nullptr; c->long_hstage_cap = 0;
                RSD_CUDA(cudaMallocHost(&c->long_hstage, stage_need + stage_need / 4));
                c->long_hstage_cap = stage_need + stage_need / 4;
            }
            unsigned char *hst = (unsigned char *)c->long_hstage;
            {
                size_t so = 0, oo = 0, q = 0;
                for (int p : batch) {
                    LongPairPlan &Q = P[p];
                    Q.da = d_seq + so; memcpy(hst + so, a[p], (size_t)Q.m); so += al256((size_t)Q.m + 64);
                    Q.db = d_seq + so; memcpy(hst + so, b[p], (size_t)Q.n); so += al256((size_t)Q.n + 64);
                    unsigned char *small = d_small + 64 * q;
                    Q.keyacc = (long long *)small; Q.dist = (double *)(small + 16); Q.state = (int *)(small + 32); Q.n_ops = (int32_t *)(small + 48);
                    if (want_script) { Q.op = d_op + oo; if (want_ij) { Q.oi = d_oi + oo; Q.oj = d_oj + oo; } oo += long_ops_slot(Q); }
                    ++q;
                }
            }
            RSD_CUDA(cudaMemcpyAsync(d_seq, hst, seq_bytes, cudaMemcpyHostToDevice, st));
            RSD_CUDA(cudaMemsetAsync(d_small, 0, small_bytes, st));                 // key accumulators 0, traceback "not started"
            for (int p : batch) {
                off += long_layout(P[p], base + off, want_script != 0, want_ij);
                P[p].bound = (unsigned long long *)(bbase + boff); boff += long_bound_bytes(P[p]);
            }
            // ---- job list of every launch of this batch ----
            std::vector<LongJob2> jobs;
            struct Launch { int job0, n_jobs; bool dirs; int ring_job0[RSD_LONG2_MAX_RINGS + 1], ring_cta0[RSD_LONG2_MAX_RINGS + 1], n_rings; int tb_pair, tb_r0; };
            std::vector<Launch> launches;
            auto next_gen = [&]() -> unsigned { if (++c->long_gen == 0u) c->long_gen = 1u; return c->long_gen; };       // (a wrap after 2^32 launch groups could meet a stale tag: ignored)
            auto make_job = [&](const LongPairPlan &Q, int64_t r0, int64_t r1, int w_lo, int w_cnt, const uint32_t *top, uint32_t *bottom, bool dirs, bool recompute, unsigned gen) {
                LongJob2 J{};
                J.gen = gen;
                J.a = Q.da; J.b = Q.db; J.m = (int)Q.m; J.n = (int)Q.n; J.r0 = (int)r0; J.r1 = (int)r1;
                J.n_panels = Q.n_panels; J.n_pad = (int)Q.n_pad; J.w_lo = w_lo; J.w_cnt = w_cnt;
                J.top = top; J.bottom = bottom; J.dirs = dirs ? Q.dirs : nullptr; J.bound = Q.bound; J.bstride = (int)Q.hb;
                J.keyacc = Q.keyacc + (recompute ? 1 : 0); J.dist = recompute ? Q.dist + 1 : Q.dist; J.S = Q.S;
                return J;
            };
            if (!solo) {
                // rings: longest pair first, each to the ring with the fewest rows so far; fewer rings when the CTAs
                // of all rings together would not be co-resident
                std::vector<int> by = batch;
                std::sort(by.begin(), by.end(), [&](int x, int y) { return P[x].m * P[x].n > P[y].m * P[y].n; });
                std::vector<std::vector<int>> ring;
                int G = (int)std::min<size_t>((size_t)rings_want, batch.size());
                for (;; --G) {
                    ring.assign((size_t)G, {});
                    std::vector<int64_t> load((size_t)G, 0);
                    for (int p : by) { int g = (int)(std::min_element(load.begin(), load.end()) - load.begin()); ring[(size_t)g].push_back(p); load[(size_t)g] += P[p].m; }
                    int64_t cta = 0;
                    for (int g = 0; g < G; ++g) { int width = 0; for (int p : ring[(size_t)g]) width = std::max(width, P[p].n_panels); cta += width; }
                    if (cta <= max_ctas || G == 1) break;
                }
                Launch L{}; L.job0 = 0; L.dirs = want_script != 0; L.n_rings = G; L.tb_pair = -1;
                const unsigned gen = next_gen();
                int cta = 0;
                for (int g = 0; g < G; ++g) {
                    L.ring_job0[g] = (int)jobs.size(); L.ring_cta0[g] = cta;
                    int width = 0;
                    for (int p : ring[(size_t)g]) { jobs.push_back(make_job(P[p], 0, P[p].m, 0, P[p].n_panels, nullptr, nullptr, want_script != 0, false, gen)); width = std::max(width, P[p].n_panels); }
                    cta += width;
                }
                L.ring_job0[G] = (int)jobs.size(); L.ring_cta0[G] = cta; L.n_jobs = (int)jobs.size();
                if (cta > max_ctas) return rsd_fail(RSD_ERANGE, "rsd_long_pairs: %d CTAs exceed the %lld co-resident ones", cta, (long long)max_ctas);
                launches.push_back(L);
            } else {
                const LongPairPlan &Q = P[batch[0]];
                auto ckpt_row = [&](int k) -> uint32_t * {          // key row of block boundary k (1 .. nb-1)
                    if (k <= 0 || k >= Q.nb) return nullptr;
                    return Q.ckpt + (size_t)(want_script ? k - 1 : (k & 1)) * (size_t)Q.n_pad;
                };
                auto add_block = [&](int blk, bool dirs, bool recompute) {
                    const int64_t r0 = (int64_t)blk * Q.hb, r1 = std::min<int64_t>(Q.m, r0 + Q.hb);
                    const unsigned gen = next_gen();                 // one tag per row block: its panel ranges read each other's columns
                    for (int r = 0; r < Q.nr; ++r) {
                        const int w_lo = (int)((int64_t)r * max_ctas), w_cnt = (int)std::min<int64_t>(max_ctas, Q.n_panels - w_lo);
                        Launch L{}; L.job0 = (int)jobs.size(); L.n_jobs = 1; L.dirs = dirs; L.n_rings = 1;
                        L.ring_job0[0] = 0; L.ring_job0[1] = 1; L.ring_cta0[0] = 0; L.ring_cta0[1] = w_cnt;
                        L.tb_pair = (dirs && r == Q.nr - 1) ? batch[0] : -1; L.tb_r0 = (int)r0;
                        jobs.push_back(make_job(Q, r0, r1, w_lo, w_cnt, ckpt_row(blk), recompute ? nullptr : ckpt_row(blk + 1), dirs, recompute, gen));
                        launches.push_back(L);
                    }
                };
                for (int blk = 0; blk < Q.nb; ++blk) add_block(blk, want_script && blk == Q.nb - 1, false);
                if (want_script) for (int blk = Q.nb - 2; blk >= 0; --blk) add_block(blk, true, true);
            }
            // traceback / emit job tables
            std::vector<LongTb2> tbs; std::vector<LongEmit2> ems;
            if (want_script) {
                if (!solo) for (int p : batch) {
                    const LongPairPlan &Q = P[p];
                    tbs.push_back(LongTb2{(int)Q.m, (int)Q.n, 0, (int)Q.n_pad, C, Q.dirs, Q.tmp, Q.state, Q.n_ops, 1});
                } else for (const Launch &L : launches) if (L.tb_pair >= 0) {
                    const LongPairPlan &Q = P[L.tb_pair];
                    tbs.push_back(LongTb2{(int)Q.m, (int)Q.n, L.tb_r0, (int)Q.n_pad, C, Q.dirs, Q.tmp, Q.state, Q.n_ops, L.tb_r0 == 0 ? 1 : 0});
                }
                for (int p : batch) { const LongPairPlan &Q = P[p]; ems.push_back(LongEmit2{Q.tmp, (int)Q.m, (int)Q.n, Q.n_ops, Q.op, Q.oi, Q.oj}); }
            }
            const size_t jb = al256(jobs.size() * sizeof(LongJob2)), tb = al256(tbs.size() * sizeof(LongTb2)), eb = al256(ems.size() * sizeof(LongEmit2));
            RSD_OK_OR_RETURN(c->long_jobs.ensure(jb + tb + eb + 256));
            unsigned char *jd = (unsigned char *)c->long_jobs.p;
            RSD_CUDA(cudaMemcpyAsync(jd, jobs.data(), jobs.size() * sizeof(LongJob2), cudaMemcpyHostToDevice, st));
            if (!tbs.empty()) RSD_CUDA(cudaMemcpyAsync(jd + jb, tbs.data(), tbs.size() * sizeof(LongTb2), cudaMemcpyHostToDevice, st));
            if (!ems.empty()) RSD_CUDA(cudaMemcpyAsync(jd + jb + tb, ems.data(), ems.size() * sizeof(LongEmit2), cudaMemcpyHostToDevice, st));
            if (c->timing && first_batch) RSD_CUDA(cudaEventRecord(c->ev0, st));
            int tb_at = 0;
            for (const Launch &L : launches) {
                LongLaunch2 LL{};
                LL.jobs = (const LongJob2 *)jd + L.job0; LL.n_rings = L.n_rings;
                for (int g = 0; g <= L.n_rings; ++g) { LL.ring_job0[g] = L.ring_job0[g]; LL.ring_cta0[g] = L.ring_cta0[g]; }
                const IntCosts *dic = c->d_ic;
                void *args[] = {&LL, &dic};
                RSD_CUDA(cudaLaunchCooperativeKernel(L.dirs ? kfn_d : kfn_n, dim3((unsigned)L.ring_cta0[L.n_rings]), dim3(32), args, 0, st));
                c->launches += 1;
                if (solo && L.tb_pair >= 0) {
                    k_long2_traceback<<<1, 32, 0, st>>>((const LongTb2 *)(jd + jb) + tb_at);
                    ++tb_at; c->launches += 1;
                }
            }
            if (c->timing) RSD_CUDA(cudaEventRecord(c->ev_t1[0], st));                 // end of the forward launches of this batch
            if (want_script) {
                if (!solo) { k_long2_traceback<<<(unsigned)tbs.size(), 32, 0, st>>>((const LongTb2 *)(jd + jb)); c->launches += 1; }
                k_long2_emit<<<(unsigned)ems.size(), 1024, 0, st>>>((const LongEmit2 *)(jd + jb + tb));
                c->launches += 1;
            }
            if (c->timing) { RSD_CUDA(cudaEventRecord(c->ev1, st)); c->timed = true; }
            RSD_CUDA(cudaGetLastError());
            // ---- results: scalars and scripts in (at most) four copies into the staging buffer, then out to the caller ----
            // (the host has to be done with the staged sequences before they are overwritten: the stream order guarantees it)
            RSD_CUDA(cudaMemcpyAsync(hst, d_small, small_bytes, cudaMemcpyDeviceToHost, st));
            unsigned char *h_op = hst + small_bytes, *h_oi = h_op + ops_entries, *h_oj = h_oi + 4 * ops_entries;
            if (want_script) {
                RSD_CUDA(cudaMemcpyAsync(h_op, d_op, ops_entries, cudaMemcpyDeviceToHost, st));
                if (want_ij) {
                    RSD_CUDA(cudaMemcpyAsync(h_oi, d_oi, 4 * ops_entries, cudaMemcpyDeviceToHost, st));
                    RSD_CUDA(cudaMemcpyAsync(h_oj, d_oj, 4 * ops_entries, cudaMemcpyDeviceToHost, st));
                }
            }
            const double t_enq = since_ms(t_entry);
            RSD_CUDA(cudaStreamSynchronize(st));
            const double t_sync = since_ms(t_entry);
            if (c->timing) { float ms = 0.f; if (cudaEventElapsedTime(&ms, c->ev0, c->ev_t1[0]) == cudaSuccess) fwd_ms_total = ms; }
            {
                size_t oo = 0;
                for (size_t q = 0; q < batch.size(); ++q) {
                    const int p = batch[q]; const LongPairPlan &Q = P[p];
                    const unsigned char *small = hst + 64 * q;
                    memcpy(&dist[p], small + 16, sizeof(double));
                    if (want_script) {
                        int32_t k32 = 0; memcpy(&k32, small + 48, sizeof k32);
                        n_ops[p] = k32;
                        memcpy(op[p], h_op + oo, (size_t)k32);
                        if (oi && oi[p]) memcpy(oi[p], h_oi + 4 * oo, sizeof(int32_t) * (size_t)k32);
                        if (oj && oj[p]) memcpy(oj[p], h_oj + 4 * oo, sizeof(int32_t) * (size_t)k32);
                        oo += long_ops_slot(Q);
                    }
                }
            }
            if (ltrace) fprintf(stderr, "[rsd trace] long batch: %zu pair(s)%s, %zu launch(es), %zu MB, blocks %d x ranges %d of pair %d; host ms since entry: enqueued %.2f, device done %.2f, results out %.2f\n",
                                batch.size(), solo ? " (blocked)" : "", launches.size(), need >> 20, P[batch[0]].nb, P[batch[0]].nr, batch[0], t_enq, t_sync, since_ms(t_entry));
            first_batch = false;
        }
        c->long_fwd_ms = fwd_ms_total;
    }
    // a pool of tens of GB (one huge pair) is not kept: the next call of any kind finds the memory free again; the pool
    // of an ordinary batch stays (cudaMalloc of 13 GB costs more than the batch itself)
    if (c->long_pool.cap > ((size_t)32 << 30)) { RSD_CUDA(cudaStreamSynchronize(st)); c->long_pool.release(); c->long_bound_hw = 0; }
    // ---- pairs that need the exact-double / fp64 kernels ----
    for (int p : fallback) {
        int mo = 0;
        RSD_OK_OR_RETURN(long_pair_v1(c, a[p], m[p], b[p], n[p], force_mode, want_script, want_script ? max_ops[p] : 0,
                                       want_script ? op[p] : nullptr, (want_script && oi) ? oi[p] : nullptr, (want_script && oj) ? oj[p] : nullptr,
                                       want_script ? &n_ops[p] : nullptr, &dist[p], &mo));
        if (mode_out) mode_out[p] = mo;
    }
    return RSD_OK;
}

extern "C" int rsd_long_pair(rsd_ctx *c, const uint8_t *a, int64_t m, const uint8_t *b, int64_t n,
                             int force_mode, int want_script, int64_t max_ops,
                             uint8_t *op, int32_t *oi, int32_t *oj, int64_t *n_ops, double *dist, int *mode_out) {
    if (getenv("RSD_LONG_V1") || getenv("RSD_LONG_R1"))
        return long_pair_v1(c, a, m, b, n, force_mode, want_script, max_ops, op, oi, oj, n_ops, dist, mode_out);
    if (!dist) return rsd_fail(RSD_EINVAL, "rsd_long_pair: bad arguments");
    if (want_script && (!op || !n_ops || max_ops < m + n)) return rsd_fail(RSD_EINVAL, "rsd_long_pair: script buffers missing or max_ops < m+n");
    return rsd_long_pairs(c, 1, &a, &m, &b, &n, force_mode, want_script, &max_ops, &op, &oi, &oj, n_ops, dist, mode_out);
}

extern "C" double rsd_long_forward_ms(rsd_ctx *c) { return c ? (double)c->long_fwd_ms : 0.0; }
