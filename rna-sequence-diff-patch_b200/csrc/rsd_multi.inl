// rsd_multi.inl — database search over several GPUs of one box from ONE process (included by rsd_api.cu).
//
// Replaces: the process fan-out of IRMethods.create_search_threads (IR:480-515: one forked process per method,
// each scanning the whole collection) on the wf_score path; north_star item (4): the sequence database is sharded
// across the GPUs, the query batch goes to every GPU, each GPU keeps a local top-k, and only the top-k lists are
// gathered with NCCL over NVLink.  One rsd_ctx per device inside one process (the reference's callers fork, so a
// process-per-GPU launcher cannot sit under search_collection), driven by one host thread per
// device; ncclCommInitAll gives one communicator per device and the gather is a grouped ncclAllGather of
// Q * k * 16 bytes per device.  NCCL is bound at run time (dlopen; the copy a host application already loaded —
// e.g. the one inside PyTorch — is reused), so librsd.so itself carries no NCCL dependency.
//
// The same device may be listed more than once (shards emulated on one GPU: the single-GPU test boxes): NCCL
// refuses duplicate devices, so the gather then degenerates to device-to-device copies on that one device.
#include <dlfcn.h>
#include <nccl.h>          // types and prototypes only; nothing is linked

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    int load() {
        if (handle) return RSD_OK;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) { handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (handle) break; }     // already in the process?
        for (const char *nm : names) { if (handle) break; handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); }
        if (!handle) return rsd_fail(RSD_ENODEV, "rsd_multi: libnccl.so.2 not found (%s); it is needed for more than one GPU", dlerror());
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllGather || !GroupStart || !GroupEnd || !GetErrorString)
            return rsd_fail(RSD_ENODEV, "rsd_multi: libnccl lacks an expected symbol");
        return RSD_OK;
    }
};
static NcclApi g_nccl;

#define RSD_NCCL(call)                                                                                          \
    do {                                                                                                        \
        ncclResult_t r__ = (call);                                                                              \
        if (r__ != ncclSuccess) return rsd_fail(RSD_ECUDA, "%s: %s", #call, g_nccl.GetErrorString(r__));        \
    } while (0)

struct rsd_multi {
    std::vector<rsd_ctx *> ctx;
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;            // empty when the gather runs without NCCL (one device, or duplicates)
    std::vector<int64_t> lo, hi;              // record range of every shard
    std::vector<DevBuf> local, gathered;      // per device: [2][Q][k] int64 (indices, fp64 score bits) and [G][2][Q][k]
    int64_t db_n = 0;
    bool loaded = false;
    pid_t pid = 0;
};

// run f(g) for every shard on its own host thread; the first failure (and its message) is reported
template <typename F>
static int multi_for_each(rsd_multi *m, F f) {
    const int G = (int)m->ctx.size();
    std::vector<int> rc((size_t)G, RSD_OK);
    std::vector<std::string> msg((size_t)G);
    auto body = [&](int g) { rc[(size_t)g] = f(g); if (rc[(size_t)g]) msg[(size_t)g] = rsd_last_error(); };
    if (G == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back(body, g);
        for (auto &t : th) t.join();
    }
    for (int g = 0; g < G; ++g) if (rc[(size_t)g]) return rsd_fail(rc[(size_t)g], "device %d: %s", m->devices[(size_t)g], msg[(size_t)g].c_str());
    return RSD_OK;
}

extern "C" int rsd_multi_create(const int *devices, int n_devices, rsd_multi **out) {
    if (!out || n_devices < 0) return rsd_fail(RSD_EINVAL, "rsd_multi_create: bad arguments");
    const int visible = rsd_device_count();
    if (visible == 0) return rsd_fail(RSD_ENODEV, "rsd: no CUDA device available; librsd has no CPU fallback");
    std::vector<int> dev;
    if (!devices || n_devices == 0) for (int d = 0; d < visible; ++d) dev.push_back(d);
    else dev.assign(devices, devices + n_devices);
    for (int d : dev) if (d < 0 || d >= visible) return rsd_fail(RSD_EINVAL, "rsd_multi_create: device %d out of range (0..%d)", d, visible - 1);
    rsd_multi *m = new (std::nothrow) rsd_multi();
    if (!m) return rsd_fail(RSD_ENOMEM, "rsd_multi_create: out of host memory");
    m->devices = dev; m->pid = getpid();
    for (int d : dev) {
        rsd_ctx *c = nullptr;
        if (int rc = rsd_create(d, &c)) { for (rsd_ctx *x : m->ctx) rsd_destroy(x); delete m; return rc; }
        m->ctx.push_back(c);
    }
    const size_t G = dev.size();
    m->lo.assign(G, 0); m->hi.assign(G, 0); m->local.resize(G); m->gathered.resize(G);
    *out = m;
    return RSD_OK;
}

extern "C" int rsd_multi_device_count(rsd_multi *m) { return m ? (int)m->ctx.size() : 0; }

extern "C" int rsd_multi_destroy(rsd_multi *m) {
    if (!m) return RSD_OK;
    if (m->pid == getpid()) {
        for (ncclComm_t c : m->comms) if (c) g_nccl.CommDestroy(c);
        for (size_t g = 0; g < m->ctx.size(); ++g) {
            if (m->ctx[g]->inited) { cudaSetDevice(m->devices[g]); m->local[g].release(); m->gathered[g].release(); }
            rsd_destroy(m->ctx[g]);
        }
    }
    delete m;
    return RSD_OK;
}

extern "C" int rsd_multi_set_costs(rsd_multi *m, double ins, double del, const double *sub) {
    if (!m) return rsd_fail(RSD_EINVAL, "rsd_multi: handle is NULL");
    for (rsd_ctx *c : m->ctx) RSD_OK_OR_RETURN(rsd_set_costs(c, ins, del, sub));
    return RSD_OK;
}

// contiguous shards balanced by the number of symbols ("balance by sum of len")
extern "C" int rsd_multi_db_load(rsd_multi *m, const uint32_t *words, const int64_t *start, const int32_t *len,
                                 int64_t n_records, int64_t n_words, int bits, uint32_t symmask) {
    if (!m) return rsd_fail(RSD_EINVAL, "rsd_multi: handle is NULL");
    if (n_records < 0 || (n_records > 0 && (!words || !start || !len))) return rsd_fail(RSD_EINVAL, "rsd_multi_db_load: bad arguments");
    if (m->pid != getpid()) return rsd_fail(RSD_EINVAL, "rsd_multi: handle used after fork()");
    const int G = (int)m->ctx.size();
    double total = 0;
    for (int64_t i = 0; i < n_records; ++i) total += len[i];
    {
        int64_t r = 0; double acc = 0;
        for (int g = 0; g < G; ++g) {
            m->lo[(size_t)g] = r;
            const double target = total * (g + 1) / G;
            if (g == G - 1) r = n_records;
            else while (r < n_records && acc + len[r] <= target) acc += len[r++];
            m->hi[(size_t)g] = r;
        }
    }
    const int per = 32 / (bits == 2 ? 2 : 4);
    RSD_OK_OR_RETURN(multi_for_each(m, [&](int g) -> int {
        const int64_t lo = m->lo[(size_t)g], hi = m->hi[(size_t)g], n = hi - lo;
        if (n == 0) return rsd_db_load(m->ctx[(size_t)g], words, start, len, 0, 0, bits, symmask, lo);
        // the shard's records may lie anywhere in the word buffer: rebase its offsets on its lowest word
        int64_t w_lo = INT64_MAX, w_hi = 0;
        for (int64_t i = lo; i < hi; ++i) { w_lo = std::min(w_lo, start[i]); w_hi = std::max(w_hi, start[i] + ((int64_t)len[i] + per - 1) / per); }
        if (w_lo < 0 || w_hi > n_words) return rsd_fail(RSD_EINVAL, "rsd_multi_db_load: a record lies outside the word buffer");
        std::vector<int64_t> st((size_t)n);
        for (int64_t i = 0; i < n; ++i) st[(size_t)i] = start[lo + i] - w_lo;
        return rsd_db_load(m->ctx[(size_t)g], words + w_lo, st.data(), len + lo, n, w_hi - w_lo, bits, symmask, lo);
    }));
    m->db_n = n_records; m->loaded = true;
    return RSD_OK;
}

extern "C" int rsd_multi_db_free(rsd_multi *m) {
    if (!m) return rsd_fail(RSD_EINVAL, "rsd_multi: handle is NULL");
    for (rsd_ctx *c : m->ctx) rsd_db_free(c);
    m->loaded = false; m->db_n = 0;
    return RSD_OK;
}

static int multi_ensure_comms(rsd_multi *m) {
    const int G = (int)m->ctx.size();
    if (G == 1 || !m->comms.empty()) return RSD_OK;
    std::vector<int> sorted = m->devices;
    std::sort(sorted.begin(), sorted.end());
    if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) return RSD_OK;       // duplicates: copies, see the header
    RSD_OK_OR_RETURN(g_nccl.load());
    m->comms.assign((size_t)G, nullptr);
    ncclResult_t r = g_nccl.CommInitAll(m->comms.data(), G, m->devices.data());
    if (r != ncclSuccess) { m->comms.clear(); return rsd_fail(RSD_ECUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(r)); }
    return RSD_OK;
}

// top_idx / top_score [n_queries][k] over the whole database; all_scores (optional) [n_queries][n_records] in record order
extern "C" int rsd_multi_db_search_topk(rsd_multi *m, const uint32_t *q_words, const int64_t *q_start, const int32_t *q_len,
                                        int64_t n_queries, int64_t q_nwords, int bits, uint32_t q_symmask, int k, int force_mode,
                                        int64_t *top_idx, double *top_score, double *all_scores, int *mode_out) {
    if (!m) return rsd_fail(RSD_EINVAL, "rsd_multi: handle is NULL");
    if (!m->loaded) return rsd_fail(RSD_EINVAL, "rsd_multi_db_search_topk: no database loaded (rsd_multi_db_load)");
    if (m->pid != getpid()) return rsd_fail(RSD_EINVAL, "rsd_multi: handle used after fork()");
    if (n_queries < 0 || (n_queries > 0 && (!q_words || !q_start || !q_len))) return rsd_fail(RSD_EINVAL, "rsd_multi_db_search_topk: bad arguments");
    if (k < 0 || k > RSD_TOPK_MAX) return rsd_fail(RSD_EINVAL, "rsd_multi_db_search_topk: k must be in 0..%d", RSD_TOPK_MAX);
    if (k > 0 && (!top_idx || !top_score)) return rsd_fail(RSD_EINVAL, "rsd_multi_db_search_topk: NULL top-k output");
    if (n_queries == 0) return RSD_OK;
    const int G = (int)m->ctx.size();
    for (rsd_ctx *c : m->ctx) RSD_OK_OR_RETURN(c->ensure_device());
    RSD_OK_OR_RETURN(multi_ensure_comms(m));
    const int kk = std::max(k, 1);
    const size_t cells = (size_t)n_queries * kk;                       // per half (indices | scores)
    int64_t max_qlen = 0;
    for (int64_t i = 0; i < n_queries; ++i) max_qlen = std::max<int64_t>(max_qlen, q_len[i]);
    std::vector<int> modes((size_t)G, 0);
    std::vector<std::vector<double>> shard_scores((size_t)(all_scores ? G : 0));
    // 1. every device: queries up, local search into its [2][Q][k] block, all asynchronous on the context's stream
    RSD_OK_OR_RETURN(multi_for_each(m, [&](int g) -> int {
        rsd_ctx *c = m->ctx[(size_t)g];
        RSD_OK_OR_RETURN(c->ensure_device());
        cudaStream_t st = c->stream;
        RSD_OK_OR_RETURN(m->local[(size_t)g].ensure(sizeof(int64_t) * 2 * cells));
        RSD_OK_OR_RETURN(m->gathered[(size_t)g].ensure(sizeof(int64_t) * 2 * cells * (size_t)G));
        RSD_OK_OR_RETURN(c->upload_seqs(c->bufQ, q_words, q_start, q_len, n_queries, q_nwords, st));
        const int64_t n_shard = m->hi[(size_t)g] - m->lo[(size_t)g];
        double *alls_dev = nullptr;
        if (all_scores && n_shard > 0) {
            RSD_OK_OR_RETURN(c->out_f64.ensure(sizeof(double) * (size_t)n_queries * (size_t)n_shard));
            alls_dev = (double *)c->out_f64.p;
        }
        int64_t *li = (int64_t *)m->local[(size_t)g].p; double *ls = (double *)(li + cells);
        RSD_OK_OR_RETURN(c->search_dev((const uint32_t *)c->bufQ.words.p, (const int64_t *)c->bufQ.start.p, (const int32_t *)c->bufQ.len.p,
                                       n_queries, max_qlen, bits, q_symmask, k, force_mode, li, ls, alls_dev, &modes[(size_t)g], st));
        if (alls_dev) {
            shard_scores[(size_t)g].resize((size_t)n_queries * (size_t)n_shard);
            RSD_CUDA(cudaMemcpyAsync(shard_scores[(size_t)g].data(), alls_dev, sizeof(double) * (size_t)n_queries * (size_t)n_shard, cudaMemcpyDeviceToHost, st));
        }
        return RSD_OK;
    }));
    if (mode_out) *mode_out = modes[0];
    // 2. the one exchange of the path: every device's top-k block to every device
    if (k > 0) {
        if (!m->comms.empty()) {
            RSD_NCCL(g_nccl.GroupStart());
            for (int g = 0; g < G; ++g) {
                ncclResult_t r = g_nccl.AllGather(m->local[(size_t)g].p, m->gathered[(size_t)g].p, 2 * cells, ncclInt64, m->comms[(size_t)g], m->ctx[(size_t)g]->stream);
                if (r != ncclSuccess) { g_nccl.GroupEnd(); return rsd_fail(RSD_ECUDA, "ncclAllGather: %s", g_nccl.GetErrorString(r)); }
            }
            RSD_NCCL(g_nccl.GroupEnd());
        } else {
            // one device (possibly listed several times): plain copies into device 0's gathered buffer
            rsd_ctx *c0 = m->ctx[0];
            RSD_OK_OR_RETURN(c0->ensure_device());
            for (int g = 0; g < G; ++g) {
                if (g > 0) { RSD_CUDA(cudaStreamSynchronize(m->ctx[(size_t)g]->stream)); }
                RSD_CUDA(cudaMemcpyAsync((int64_t *)m->gathered[0].p + (size_t)g * 2 * cells, m->local[(size_t)g].p, sizeof(int64_t) * 2 * cells,
                                         cudaMemcpyDeviceToDevice, c0->stream));
            }
        }
    }
    // 3. device 0's copy of the gathered lists comes back; merge with the same key (score desc, index asc)
    std::vector<int64_t> host((size_t)(k > 0 ? 2 * cells * (size_t)G : 0));
    {
        rsd_ctx *c0 = m->ctx[0];
        RSD_OK_OR_RETURN(c0->ensure_device());
        if (k > 0) RSD_CUDA(cudaMemcpyAsync(host.data(), m->gathered[0].p, sizeof(int64_t) * host.size(), cudaMemcpyDeviceToHost, c0->stream));
    }
    for (int g = 0; g < G; ++g) { RSD_OK_OR_RETURN(m->ctx[(size_t)g]->ensure_device()); RSD_CUDA(cudaStreamSynchronize(m->ctx[(size_t)g]->stream)); }
    if (k > 0) {
        std::vector<int64_t> gi((size_t)G * cells); std::vector<double> gs((size_t)G * cells);
        for (int g = 0; g < G; ++g) {
            memcpy(gi.data() + (size_t)g * cells, host.data() + (size_t)g * 2 * cells, sizeof(int64_t) * cells);
            memcpy(gs.data() + (size_t)g * cells, host.data() + (size_t)g * 2 * cells + cells, sizeof(double) * cells);
        }
        RSD_OK_OR_RETURN(rsd_topk_merge(gi.data(), gs.data(), G, n_queries, k, top_idx, top_score));
    }
    if (all_scores)
        for (int g = 0; g < G; ++g) {
            const int64_t lo = m->lo[(size_t)g], n_shard = m->hi[(size_t)g] - lo;
            for (int64_t q = 0; q < n_queries && n_shard > 0; ++q)
                memcpy(all_scores + (size_t)q * (size_t)m->db_n + lo, shard_scores[(size_t)g].data() + (size_t)q * (size_t)n_shard, sizeof(double) * (size_t)n_shard);
        }
    return RSD_OK;
}

extern "C" int64_t rsd_multi_launch_count(rsd_multi *m) {
    int64_t n = 0;
    if (m) for (rsd_ctx *c : m->ctx) n += c->launches;
    return n;
}
