// api.cu -- extern "C" entry points declared in include/vr180_b200.h (device-pointer API + host-buffer pipeline).
#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace vr180 {

std::atomic<uint64_t> g_launches{0};
static thread_local std::string t_cuda_error;

void set_cuda_error(cudaError_t e, const char* where) {
    t_cuda_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + where;
    cudaGetLastError();  // clear the sticky-less error state
}

const std::vector<int16_t>& host_table(int K);  // kernels.cu

}  // namespace vr180

using namespace vr180;

extern "C" {

int vr180_abi_version(void) { return VR180_ABI_VERSION; }

const char* vr180_status_string(int status) {
    switch (status) {
        case VR180_OK: return "ok";
        case VR180_ERR_INVALID_ARG: return "invalid argument";
        case VR180_ERR_UNSUPPORTED: return "unsupported configuration";
        case VR180_ERR_CUDA: return "CUDA error";
        case VR180_ERR_NO_DEVICE: return "no CUDA device";
        case VR180_ERR_CHAIN: return "malformed chain descriptor";
        case VR180_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

const char* vr180_last_cuda_error(void) { return t_cuda_error.c_str(); }

int vr180_device_info(int device, char* name, size_t name_len, int* sm_count, int* cc_major, int* cc_minor) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (device >= 0 && device < n) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
            if (name && name_len) {
                strncpy(name, prop.name, name_len - 1);
                name[name_len - 1] = 0;
            }
            if (sm_count) *sm_count = prop.multiProcessorCount;
            if (cc_major) *cc_major = prop.major;
            if (cc_minor) *cc_minor = prop.minor;
        }
    }
    return n;
}

uint64_t vr180_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

/* test hook: the host-built OpenCV weight tables (K = 4 cubic, K = 8 lanczos4), 1024*K*K int16 */
int vr180_debug_weight_table(int K, int16_t* out) {
    if ((K != 4 && K != 8) || !out) return VR180_ERR_INVALID_ARG;
    const auto& t = host_table(K);
    memcpy(out, t.data(), t.size() * sizeof(int16_t));
    return VR180_OK;
}

int vr180_build_map(const vr180_chain_t* chain, int out_w, int out_h, float* xmap_dev, float* ymap_dev,
                    int64_t map_pitch, void* stream) {
    if (!chain || !xmap_dev || !ymap_dev || out_w <= 0 || out_h <= 0 || map_pitch < out_w) return VR180_ERR_INVALID_ARG;
    const int rc = validate_chain(chain);
    if (rc != VR180_OK) return rc;
    DeviceGuard g(xmap_dev);
    if (!g.ok) return VR180_ERR_INVALID_ARG;
    return launch_build_map(chain, out_w, out_h, xmap_dev, ymap_dev, map_pitch, (cudaStream_t)stream);
}

int vr180_pack_lut(const float* xmap_dev, const float* ymap_dev, int64_t map_pitch, int out_w, int out_h,
                   int32_t* fixed_dev, int64_t fixed_pitch, void* stream) {
    if (!xmap_dev || !ymap_dev || !fixed_dev || out_w <= 0 || out_h <= 0 || map_pitch < out_w || fixed_pitch < out_w)
        return VR180_ERR_INVALID_ARG;
    DeviceGuard g(fixed_dev);
    if (!g.ok) return VR180_ERR_INVALID_ARG;
    return launch_pack_lut(xmap_dev, ymap_dev, map_pitch, out_w, out_h, fixed_dev, fixed_pitch, (cudaStream_t)stream);
}

int vr180_remap(const vr180_remap_params_t* params, void* stream) {
    if (!params || !params->dst) return VR180_ERR_INVALID_ARG;
    DeviceGuard g(params->dst);
    if (!g.ok) return VR180_ERR_INVALID_ARG;
    return launch_remap(params, (cudaStream_t)stream);
}

int vr180_get_radius(const vr180_image_t* views, int n_views, int n_frames, double threshold, int32_t* transitions_dev,
                     double* radius_dev, void* stream) {
    if (!views || (!transitions_dev && !radius_dev)) return VR180_ERR_INVALID_ARG;
    DeviceGuard g(transitions_dev ? (const void*)transitions_dev : (const void*)radius_dev);
    if (!g.ok) return VR180_ERR_INVALID_ARG;
    return launch_get_radius(views, n_views, n_frames, threshold, transitions_dev, radius_dev, (cudaStream_t)stream);
}

/* ---------------------------------------------------------------------------------------------------------
 * pinned host memory helpers
 * ------------------------------------------------------------------------------------------------------- */
int vr180_host_alloc(size_t bytes, void** out) {
    if (!out) return VR180_ERR_INVALID_ARG;
    VR180_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return VR180_OK;
}
int vr180_host_free(void* p) {
    if (p) VR180_CUDA(cudaFreeHost(p));
    return VR180_OK;
}
int vr180_host_register(void* p, size_t bytes) {
    if (!p) return VR180_ERR_INVALID_ARG;
    VR180_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return VR180_OK;
}
int vr180_host_unregister(void* p) {
    if (!p) return VR180_ERR_INVALID_ARG;
    VR180_CUDA(cudaHostUnregister(p));
    return VR180_OK;
}

}  // extern "C"

/* ---------------------------------------------------------------------------------------------------------
 * host-buffer pipeline
 * ------------------------------------------------------------------------------------------------------- */
namespace {

constexpr int kSlots = 2;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return VR180_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_cuda_error(e, "cudaMalloc");
            return e == cudaErrorMemoryAllocation ? VR180_ERR_NOMEM : VR180_ERR_CUDA;
        }
        cap = want;
        return VR180_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct vr180_ctx {
    int device = 0;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[kSlots] = {}, ev_comp[kSlots] = {}, ev_d2h[kSlots] = {};
    DevBuf src[kSlots][2], dst[kSlots], maps[2][2], trans, radius;
    uint64_t map_key = 0;
    size_t map_elems = 0;
    std::mutex mu;
};

extern "C" {

int vr180_ctx_create(int device, vr180_ctx_t** out) {
    if (!out) return VR180_ERR_INVALID_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return VR180_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) return VR180_ERR_INVALID_ARG;
    DeviceGuard g(device);
    if (!g.ok) return VR180_ERR_CUDA;
    vr180_ctx* c = new (std::nothrow) vr180_ctx();
    if (!c) return VR180_ERR_NOMEM;
    c->device = device;
    VR180_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    VR180_CUDA(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    VR180_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
        VR180_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        VR180_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
        VR180_CUDA(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
    }
    *out = c;
    return VR180_OK;
}

int vr180_ctx_destroy(vr180_ctx_t* c) {
    if (!c) return VR180_OK;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kSlots; ++i) {
        for (int v = 0; v < 2; ++v) c->src[i][v].release();
        c->dst[i].release();
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    for (int v = 0; v < 2; ++v)
        for (int k = 0; k < 2; ++k) c->maps[v][k].release();
    c->trans.release();
    c->radius.release();
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    delete c;
    return VR180_OK;
}

int vr180_ctx_device(const vr180_ctx_t* c) { return c ? c->device : -1; }

int vr180_ctx_run(vr180_ctx_t* c, const vr180_host_job_t* job) {
    if (!c || !job || !job->dst) return VR180_ERR_INVALID_ARG;
    const int V = job->n_views, F = job->n_frames, C = job->channels;
    if (V < 1 || V > 2 || F < 0 || job->src_rows <= 0 || job->src_cols <= 0 || job->out_w <= 0 || job->out_h <= 0)
        return VR180_ERR_INVALID_ARG;
    if (C != 1 && C != 3 && C != 4) return VR180_ERR_UNSUPPORTED;
    for (int v = 0; v < V; ++v)
        if (!job->src[v]) return VR180_ERR_INVALID_ARG;
    if (job->map_kind != VR180_MAPSRC_ANALYTIC && job->map_kind != VR180_MAPSRC_FLOAT2) return VR180_ERR_UNSUPPORTED;
    if (job->radius_mode == 1 && job->map_kind != VR180_MAPSRC_ANALYTIC) return VR180_ERR_UNSUPPORTED;
    if (F == 0) return VR180_OK;

    std::lock_guard<std::mutex> lock(c->mu);
    DeviceGuard g(c->device);
    if (!g.ok) return VR180_ERR_CUDA;

    const size_t src_row = (size_t)job->src_cols * C, src_pitch = align_up(src_row, 16);
    const size_t src_frame = src_pitch * job->src_rows;
    const size_t dst_row = (size_t)job->out_w * V * C, dst_pitch = align_up(dst_row, 16);
    const size_t dst_frame = dst_pitch * job->out_h;

    // frames per chunk: ~256 MiB of traffic per chunk so copies of neighbouring chunks overlap compute
    const size_t per_frame = src_frame * V + dst_frame;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)F, ((size_t)256 << 20) / std::max<size_t>(per_frame, 1)));
    if (F >= 2) chunk = std::min(chunk, (F + 1) / 2);
    const int n_chunks = (F + chunk - 1) / chunk;

    int rc;
    for (int s = 0; s < std::min(kSlots, n_chunks); ++s) {
        for (int v = 0; v < V; ++v)
            if ((rc = c->src[s][v].reserve(src_frame * chunk)) != VR180_OK) return rc;
        if ((rc = c->dst[s].reserve(dst_frame * chunk)) != VR180_OK) return rc;
    }
    if ((rc = c->trans.reserve(sizeof(int32_t) * 2 * V * F)) != VR180_OK) return rc;
    if ((rc = c->radius.reserve(sizeof(double) * F)) != VR180_OK) return rc;

    // maps (FLOAT2): upload once, keep while the caller's cache key is unchanged
    const int n_maps = (V == 2 && !job->share_map) ? 2 : 1;
    if (job->map_kind == VR180_MAPSRC_FLOAT2) {
        const size_t elems = (size_t)job->out_w * job->out_h;
        const bool cached = job->maps_cache_key != 0 && job->maps_cache_key == c->map_key && c->map_elems == elems;
        if (!cached) {
            for (int m = 0; m < n_maps; ++m) {
                if (!job->xmap[m] || !job->ymap[m]) return VR180_ERR_INVALID_ARG;
                if ((rc = c->maps[m][0].reserve(elems * 4)) != VR180_OK) return rc;
                if ((rc = c->maps[m][1].reserve(elems * 4)) != VR180_OK) return rc;
                VR180_CUDA(cudaMemcpyAsync(c->maps[m][0].p, job->xmap[m], elems * 4, cudaMemcpyHostToDevice, c->s_comp));
                VR180_CUDA(cudaMemcpyAsync(c->maps[m][1].p, job->ymap[m], elems * 4, cudaMemcpyHostToDevice, c->s_comp));
            }
            c->map_key = job->maps_cache_key;
            c->map_elems = elems;
        }
    } else {
        for (int m = 0; m < n_maps; ++m) {
            rc = validate_chain(job->chain[m]);
            if (rc != VR180_OK) return rc;
        }
    }

    for (int k = 0; k < n_chunks; ++k) {
        const int s = k % kSlots, f0 = k * chunk, nf = std::min(chunk, F - f0);
        // --- upload -------------------------------------------------------------------------------------
        if (k >= kSlots) VR180_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[s], 0));  // slot's previous compute done
        for (int v = 0; v < V; ++v) {
            const uint8_t* hp = job->src[v] + (size_t)f0 * job->src_frame_stride[v];
            uint8_t* dp = (uint8_t*)c->src[s][v].p;
            if (job->src_frame_stride[v] == job->src_pitch[v] * job->src_rows) {
                VR180_CUDA(cudaMemcpy2DAsync(dp, src_pitch, hp, job->src_pitch[v], src_row, (size_t)job->src_rows * nf,
                                             cudaMemcpyHostToDevice, c->s_h2d));
            } else {
                for (int f = 0; f < nf; ++f)
                    VR180_CUDA(cudaMemcpy2DAsync(dp + f * src_frame, src_pitch, hp + (size_t)f * job->src_frame_stride[v],
                                                 job->src_pitch[v], src_row, job->src_rows, cudaMemcpyHostToDevice,
                                                 c->s_h2d));
            }
        }
        VR180_CUDA(cudaEventRecord(c->ev_h2d[s], c->s_h2d));
        // --- compute ------------------------------------------------------------------------------------
        VR180_CUDA(cudaStreamWaitEvent(c->s_comp, c->ev_h2d[s], 0));
        if (k >= kSlots) VR180_CUDA(cudaStreamWaitEvent(c->s_comp, c->ev_d2h[s], 0));  // slot's previous download done
        vr180_remap_params_t p;
        memset(&p, 0, sizeof(p));
        p.n_views = V;
        p.n_frames = nf;
        p.share_map = (V == 2 && job->share_map) ? 1 : 0;
        p.out_w = job->out_w;
        p.out_h = job->out_h;
        p.interpolation = job->interpolation;
        p.border_mode = job->border_mode;
        memcpy(p.border_value, job->border_value, 4);
        p.dst = (uint8_t*)c->dst[s].p;
        p.dst_pitch = (int64_t)dst_pitch;
        p.dst_frame_stride = (int64_t)dst_frame;
        double* rad = (double*)c->radius.p + f0;
        for (int v = 0; v < V; ++v) {
            vr180_view_t& vw = p.view[v];
            vw.src.data = (const uint8_t*)c->src[s][v].p;
            vw.src.rows = job->src_rows;
            vw.src.cols = job->src_cols;
            vw.src.channels = C;
            vw.src.pitch = (int64_t)src_pitch;
            vw.src.frame_stride = (int64_t)src_frame;
            vw.dst_x_offset = v * job->out_w;
            const int m = (n_maps == 2) ? v : 0;
            vw.map.kind = job->map_kind;
            if (job->map_kind == VR180_MAPSRC_ANALYTIC) {
                vw.map.chain = job->chain[m];
                vw.map.radius_dev = job->radius_mode == 1 ? rad : nullptr;
            } else {
                vw.map.xmap = (const float*)c->maps[m][0].p;
                vw.map.ymap = (const float*)c->maps[m][1].p;
                vw.map.map_pitch = job->out_w;
            }
        }
        if (job->radius_mode == 1 || job->transitions_out) {
            vr180_image_t im[2] = {p.view[0].src, p.view[1].src};
            rc = launch_get_radius(im, V, nf, job->threshold, (int32_t*)c->trans.p + (size_t)2 * V * f0, rad, c->s_comp);
            if (rc != VR180_OK) return rc;
        }
        rc = launch_remap(&p, c->s_comp);
        if (rc != VR180_OK) return rc;
        VR180_CUDA(cudaEventRecord(c->ev_comp[s], c->s_comp));
        // --- download -----------------------------------------------------------------------------------
        VR180_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[s], 0));
        uint8_t* hd = job->dst + (size_t)f0 * job->dst_frame_stride;
        if (job->dst_frame_stride == job->dst_pitch * job->out_h) {
            VR180_CUDA(cudaMemcpy2DAsync(hd, job->dst_pitch, c->dst[s].p, dst_pitch, dst_row, (size_t)job->out_h * nf,
                                         cudaMemcpyDeviceToHost, c->s_d2h));
        } else {
            for (int f = 0; f < nf; ++f)
                VR180_CUDA(cudaMemcpy2DAsync(hd + (size_t)f * job->dst_frame_stride, job->dst_pitch,
                                             (uint8_t*)c->dst[s].p + f * dst_frame, dst_pitch, dst_row, job->out_h,
                                             cudaMemcpyDeviceToHost, c->s_d2h));
        }
        VR180_CUDA(cudaEventRecord(c->ev_d2h[s], c->s_d2h));
    }
    if (job->transitions_out || job->radius_out) {
        VR180_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[(n_chunks - 1) % kSlots], 0));
        if (job->transitions_out)
            VR180_CUDA(cudaMemcpyAsync(job->transitions_out, c->trans.p, sizeof(int32_t) * 2 * V * F,
                                       cudaMemcpyDeviceToHost, c->s_d2h));
        if (job->radius_out && job->radius_mode == 1)
            VR180_CUDA(cudaMemcpyAsync(job->radius_out, c->radius.p, sizeof(double) * F, cudaMemcpyDeviceToHost, c->s_d2h));
    }
    VR180_CUDA(cudaStreamSynchronize(c->s_d2h));
    VR180_CUDA(cudaStreamSynchronize(c->s_comp));
    return VR180_OK;
}

}  // extern "C"
