// libclane_b200.so -- host side of the C-ABI: graph build, row schedule, host-buffer session.
//
// Reference call sites replaced (all under /root/reference/clane/):
//   graph.py:104-110  Graph.A (coalesce)              -> clane_csr_from_edges
//   graph.py:130-138  Graph.Z / set_Z                  -> clane_session_get_z / set_z
//   embedder.py:71-108 Embedder.propagate              -> clane_session_propagate
//   embedder.py:56-69  Embedder.iterate                -> clane_session_iterate
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <numeric>
#include <vector>

#include "common.cuh"

extern "C" {

int64_t clane_csr_from_edges(const int64_t* h_src, const int64_t* h_dst, int64_t e_raw, int64_t n, int32_t* h_rowptr,
                             int32_t* h_col) {
    if (e_raw < 0 || n < 0 || !h_rowptr || (e_raw > 0 && (!h_src || !h_dst || !h_col))) return CLANE_EINVAL;
    if (n > INT32_MAX || e_raw > INT32_MAX) return CLANE_ERANGE;
    try {
        // counting sort by source, then sort + unique each row's destinations: the order
        // torch's coalesce() produces (row-major, ascending column, duplicates merged).
        std::vector<int64_t> start((size_t)n + 1, 0);
        for (int64_t e = 0; e < e_raw; ++e) {
            const int64_t s = h_src[e], t = h_dst[e];
            if (s < 0 || s >= n || t < 0 || t >= n) return CLANE_ERANGE;
            start[(size_t)s + 1]++;
        }
        for (int64_t v = 0; v < n; ++v) start[(size_t)v + 1] += start[(size_t)v];
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        std::vector<int32_t> tmp((size_t)std::max<int64_t>(e_raw, 1));
        for (int64_t e = 0; e < e_raw; ++e) tmp[(size_t)fill[(size_t)h_src[e]]++] = (int32_t)h_dst[e];
        int64_t out = 0;
        h_rowptr[0] = 0;
        for (int64_t v = 0; v < n; ++v) {
            int32_t* a = tmp.data() + start[(size_t)v];
            int32_t* b = tmp.data() + start[(size_t)v + 1];
            if (b - a > 1) std::sort(a, b);
            for (int32_t* p = a; p < b; ++p)
                if (p == a || *p != *(p - 1)) h_col[out++] = *p;
            h_rowptr[v + 1] = (int32_t)out;
        }
        return out;
    } catch (const std::bad_alloc&) {
        return (int64_t)cudaErrorMemoryAllocation;
    }
}

int clane_row_schedule(const int32_t* h_rowptr, int32_t n, int32_t hub_threshold, int32_t* h_light_order,
                       int32_t* n_light, int32_t* h_hub_rows, int32_t* n_hub) {
    if (!h_rowptr || n < 0 || !h_light_order || !n_light || !h_hub_rows || !n_hub || hub_threshold < 1)
        return CLANE_EINVAL;
    std::vector<int32_t> hubs, medium, rest;
    for (int32_t v = 0; v < n; ++v) {
        const int32_t k = h_rowptr[v + 1] - h_rowptr[v];
        if (k == 0) continue;
        if (k > hub_threshold) hubs.push_back(v);
        else if (k > 32) medium.push_back(v);
        else rest.push_back(v);
    }
    auto by_degree_desc = [&](int32_t x, int32_t y) {
        const int32_t kx = h_rowptr[x + 1] - h_rowptr[x], ky = h_rowptr[y + 1] - h_rowptr[y];
        return kx != ky ? kx > ky : x < y;
    };
    std::sort(hubs.begin(), hubs.end(), by_degree_desc);
    std::sort(medium.begin(), medium.end(), by_degree_desc);
    std::copy(hubs.begin(), hubs.end(), h_hub_rows);
    std::copy(medium.begin(), medium.end(), h_light_order);
    std::copy(rest.begin(), rest.end(), h_light_order + medium.size());
    *n_hub = (int32_t)hubs.size();
    *n_light = (int32_t)(medium.size() + rest.size());
    return CLANE_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer session
// ---------------------------------------------------------------------------------------------
struct clane_session {
    int32_t n = 0, d = 0, ld = 0;
    int64_t e = 0;
    int32_t n_light = 0, n_hub = 0;
    int32_t *rowptr = nullptr, *col = nullptr, *erow = nullptr, *light = nullptr, *hubs = nullptr;
    float *X = nullptr, *Z[2] = {nullptr, nullptr}, *prev = nullptr, *w = nullptr, *norms2 = nullptr;
    float *amount = nullptr, *log = nullptr;
    clane_patience* state = nullptr;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int cur = 0;  // Z[cur] holds the current embeddings
    int32_t log_cap = 0;
    cudaStream_t stream = nullptr;
    clane_patience* h_state = nullptr;  // pinned
    float* h_stage = nullptr;           // pinned staging for padded rows
    bool p_valid = false;
};

static int copy_rows_h2d(clane_session* s, float* d_dst, const float* h_src) {
    if (s->ld == s->d) {
        CLANE_CUDA(cudaMemcpyAsync(d_dst, h_src, (size_t)s->n * s->d * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    } else {
        CLANE_CUDA(cudaMemsetAsync(d_dst, 0, (size_t)s->n * s->ld * sizeof(float), s->stream));
        CLANE_CUDA(cudaMemcpy2DAsync(d_dst, (size_t)s->ld * sizeof(float), h_src, (size_t)s->d * sizeof(float),
                                     (size_t)s->d * sizeof(float), (size_t)s->n, cudaMemcpyHostToDevice, s->stream));
    }
    return CLANE_OK;
}

static int copy_rows_d2h(clane_session* s, float* h_dst, const float* d_src) {
    CLANE_CUDA(cudaMemcpy2DAsync(h_dst, (size_t)s->d * sizeof(float), d_src, (size_t)s->ld * sizeof(float),
                                 (size_t)s->d * sizeof(float), (size_t)s->n, cudaMemcpyDeviceToHost, s->stream));
    return CLANE_OK;
}

int clane_session_destroy(clane_session* s) {
    if (!s) return CLANE_OK;
    cudaFree(s->rowptr); cudaFree(s->col); cudaFree(s->erow); cudaFree(s->light); cudaFree(s->hubs);
    cudaFree(s->X); cudaFree(s->Z[0]); cudaFree(s->Z[1]); cudaFree(s->prev); cudaFree(s->w); cudaFree(s->norms2);
    cudaFree(s->amount); cudaFree(s->log); cudaFree(s->state); cudaFree(s->ws);
    if (s->h_state) cudaFreeHost(s->h_state);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return CLANE_OK;
}

#define SESSION_TRY(x)                                   \
    do {                                                 \
        int rc__ = (x);                                  \
        if (rc__ != CLANE_OK) { clane_session_destroy(s); return rc__; } \
    } while (0)
#define SESSION_CUDA(x) SESSION_TRY((int)(x))

int clane_session_create(clane_session** out, int32_t n, int64_t e, int32_t d, const int32_t* h_rowptr,
                         const int32_t* h_col, const float* h_X, int32_t hub_threshold) {
    if (!out || n < 0 || e < 0 || d < 1 || !h_rowptr || (e > 0 && !h_col) || (n > 0 && !h_X)) return CLANE_EINVAL;
    if (h_rowptr[n] != e) return CLANE_EINVAL;
    clane_session* s = new (std::nothrow) clane_session();
    if (!s) return (int)cudaErrorMemoryAllocation;
    s->n = n; s->e = e; s->d = d; s->ld = clane_padded_ld(d);
    s->log_cap = 1 << 16;
    const size_t zbytes = std::max<size_t>((size_t)n * s->ld * sizeof(float), 16);
    const size_t ebytes = std::max<size_t>((size_t)e * sizeof(int32_t), 16);
    SESSION_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    SESSION_CUDA(cudaMalloc(&s->rowptr, (size_t)(n + 1) * sizeof(int32_t)));
    SESSION_CUDA(cudaMalloc(&s->col, ebytes));
    SESSION_CUDA(cudaMalloc(&s->erow, ebytes));
    SESSION_CUDA(cudaMalloc(&s->light, std::max<size_t>((size_t)n * sizeof(int32_t), 16)));
    SESSION_CUDA(cudaMalloc(&s->hubs, std::max<size_t>((size_t)n * sizeof(int32_t), 16)));
    SESSION_CUDA(cudaMalloc(&s->X, zbytes));
    SESSION_CUDA(cudaMalloc(&s->Z[0], zbytes));
    SESSION_CUDA(cudaMalloc(&s->Z[1], zbytes));
    SESSION_CUDA(cudaMalloc(&s->prev, zbytes));
    SESSION_CUDA(cudaMalloc(&s->w, ebytes));
    SESSION_CUDA(cudaMalloc(&s->norms2, 2 * sizeof(float)));
    SESSION_CUDA(cudaMalloc(&s->amount, sizeof(float)));
    SESSION_CUDA(cudaMalloc(&s->log, (size_t)s->log_cap * sizeof(float)));
    SESSION_CUDA(cudaMalloc(&s->state, sizeof(clane_patience)));
    s->ws_bytes = clane_workspace_bytes(n, e, d);
    SESSION_CUDA(cudaMalloc(&s->ws, s->ws_bytes));
    SESSION_CUDA(cudaMallocHost(&s->h_state, sizeof(clane_patience)));

    std::vector<int32_t> light((size_t)std::max(n, 1)), hubs((size_t)std::max(n, 1));
    SESSION_TRY(clane_row_schedule(h_rowptr, n, hub_threshold > 0 ? hub_threshold : 256, light.data(), &s->n_light,
                                   hubs.data(), &s->n_hub));
    SESSION_CUDA(cudaMemcpyAsync(s->rowptr, h_rowptr, (size_t)(n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    if (e > 0) SESSION_CUDA(cudaMemcpyAsync(s->col, h_col, (size_t)e * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    if (s->n_light > 0)
        SESSION_CUDA(cudaMemcpyAsync(s->light, light.data(), (size_t)s->n_light * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    if (s->n_hub > 0)
        SESSION_CUDA(cudaMemcpyAsync(s->hubs, hubs.data(), (size_t)s->n_hub * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    if (n > 0) {
        SESSION_TRY(copy_rows_h2d(s, s->X, h_X));
        SESSION_CUDA(cudaMemcpyAsync(s->Z[0], s->X, zbytes, cudaMemcpyDeviceToDevice, s->stream));
        SESSION_CUDA(cudaMemcpyAsync(s->Z[1], s->X, zbytes, cudaMemcpyDeviceToDevice, s->stream));
    }
    SESSION_TRY(clane_edge_rows(s->rowptr, n, e, s->erow, s->stream));
    SESSION_CUDA(cudaStreamSynchronize(s->stream));  // the host vectors above go out of scope
    *out = s;
    return CLANE_OK;
}

int clane_session_set_z(clane_session* s, const float* h_Z) {
    if (!s || !h_Z) return CLANE_EINVAL;
    int rc = copy_rows_h2d(s, s->Z[0], h_Z);
    if (rc != CLANE_OK) return rc;
    CLANE_CUDA(cudaMemcpyAsync(s->Z[1], s->Z[0], (size_t)s->n * s->ld * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    s->cur = 0;
    s->p_valid = false;
    CLANE_CUDA(cudaStreamSynchronize(s->stream));
    return CLANE_OK;
}

int clane_session_get_z(clane_session* s, float* h_Z) {
    if (!s || !h_Z) return CLANE_EINVAL;
    if (s->n == 0) return CLANE_OK;
    int rc = copy_rows_d2h(s, h_Z, s->Z[s->cur]);
    if (rc != CLANE_OK) return rc;
    CLANE_CUDA(cudaStreamSynchronize(s->stream));
    return CLANE_OK;
}

static int session_build_p(clane_session* s) {
    int rc = clane_build_p_cosine(s->Z[s->cur], s->ld, s->d, s->n, s->e, s->rowptr, s->erow, s->col, s->w, s->norms2,
                                  s->ws, s->ws_bytes, s->stream);
    if (rc == CLANE_OK) s->p_valid = true;
    return rc;
}

int clane_session_build_p(clane_session* s, float* h_w) {
    if (!s) return CLANE_EINVAL;
    int rc = session_build_p(s);
    if (rc != CLANE_OK) return rc;
    if (h_w && s->e > 0)
        CLANE_CUDA(cudaMemcpyAsync(h_w, s->w, (size_t)s->e * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CLANE_CUDA(cudaStreamSynchronize(s->stream));
    return CLANE_OK;
}

static int session_sweep(clane_session* s, float gamma, bool with_state, float* d_amount) {
    int rc = clane_sweep(s->X, s->Z[s->cur], s->Z[s->cur ^ 1], s->ld, s->d, s->n, s->rowptr, s->col, s->w, gamma,
                         s->light, s->n_light, s->hubs, s->n_hub, d_amount, with_state ? s->state : nullptr,
                         with_state ? s->log : nullptr, s->log_cap, s->ws, s->ws_bytes, s->stream);
    s->cur ^= 1;
    return rc;
}

int clane_session_propagate(clane_session* s, float gamma, int32_t tol, int32_t max_sweeps, float* h_amounts,
                            int32_t cap, int32_t* sweeps) {
    if (!s || tol < 1 || max_sweeps < 0) return CLANE_EINVAL;
    int rc = session_build_p(s);
    if (rc != CLANE_OK) return rc;
    rc = clane_patience_reset(s->state, tol, max_sweeps, s->stream);
    if (rc != CLANE_OK) return rc;
    const int start = s->cur;
    int enq = 0;  // sweeps enqueued so far (the ones after the stop are device-side no-ops)
    const int batch = std::max(1, std::min(tol, 8));
    for (;;) {
        for (int i = 0; i < batch; ++i, ++enq) {
            rc = session_sweep(s, gamma, true, nullptr);
            if (rc != CLANE_OK) return rc;
        }
        CLANE_CUDA(cudaMemcpyAsync(s->h_state, s->state, sizeof(clane_patience), cudaMemcpyDeviceToHost, s->stream));
        CLANE_CUDA(cudaStreamSynchronize(s->stream));
        if (s->h_state->stop) break;
    }
    const int done = s->h_state->sweeps;
    s->cur = (start + done) & 1;  // sweep i read Z[(start+i)&1] and wrote the other buffer
    if (sweeps) *sweeps = done;
    if (h_amounts && cap > 0) {
        const int cnt = std::min(std::min(done, cap), s->log_cap);
        CLANE_CUDA(cudaMemcpyAsync(h_amounts, s->log, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        CLANE_CUDA(cudaStreamSynchronize(s->stream));
    }
    (void)enq;
    return CLANE_OK;
}

int clane_session_iterate(clane_session* s, float gamma, int32_t tol, int32_t max_outer, float* min_amount,
                          int32_t* h_sweeps_per_call, int32_t cap, int32_t* outer) {
    if (!s || tol < 1 || !min_amount) return CLANE_EINVAL;
    int patience = tol, calls = 0;
    const size_t zbytes = (size_t)s->n * s->ld * sizeof(float);
    for (;;) {
        CLANE_CUDA(cudaMemcpyAsync(s->prev, s->Z[s->cur], zbytes, cudaMemcpyDeviceToDevice, s->stream));
        int sweeps = 0;
        int rc = clane_session_propagate(s, gamma, tol, 0, nullptr, 0, &sweeps);
        if (rc != CLANE_OK) return rc;
        rc = clane_l1_diff(s->Z[s->cur], s->prev, s->ld, s->d, s->n, s->amount, s->ws, s->ws_bytes, s->stream);
        if (rc != CLANE_OK) return rc;
        float amt = 0.0f;
        CLANE_CUDA(cudaMemcpyAsync(&amt, s->amount, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        CLANE_CUDA(cudaStreamSynchronize(s->stream));
        if (h_sweeps_per_call && calls < cap) h_sweeps_per_call[calls] = sweeps;
        ++calls;
        if (*min_amount > amt) { patience = tol; *min_amount = amt; }
        else patience -= 1;
        if (patience == 0) break;
        if (max_outer > 0 && calls >= max_outer) break;
    }
    if (outer) *outer = calls;
    return CLANE_OK;
}

int clane_session_sweeps(clane_session* s, float gamma, int32_t sweeps, float* h_amount) {
    if (!s || sweeps < 0) return CLANE_EINVAL;
    if (!s->p_valid) {
        int rc = session_build_p(s);
        if (rc != CLANE_OK) return rc;
    }
    for (int i = 0; i < sweeps; ++i) {
        int rc = session_sweep(s, gamma, false, s->amount);
        if (rc != CLANE_OK) return rc;
    }
    if (h_amount) CLANE_CUDA(cudaMemcpyAsync(h_amount, s->amount, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CLANE_CUDA(cudaStreamSynchronize(s->stream));
    return CLANE_OK;
}

}  // extern "C"
