// codec.cu -- optional JPEG decode / encode on the device (nvJPEG), so that a JPEG -> warp -> JPEG conversion moves only
// compressed bytes over PCIe.  Stands in for cv.imread / cv.imwrite around the hot path
// (/root/reference/src/vr180_convert/remapper.py:373, :453, :519), OPT-IN only: nvJPEG's decoder is not bit-compatible
// with the libjpeg-turbo decoder inside cv.imread (IDCT and chroma upsampling differ by a few grey levels), so the
// default file path of apply / apply_lr keeps cv2 and its bit-exact contract.
//
// libnvjpeg is loaded with dlopen at first use (the shared object has no link-time dependency on it); when it is absent
// every entry point returns VR180_ERR_UNSUPPORTED.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

using namespace vr180;

namespace {

struct Api {
    void* so = nullptr;
    decltype(&nvjpegCreateSimple) CreateSimple = nullptr;
    decltype(&nvjpegJpegStateCreate) JpegStateCreate = nullptr;
    decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
    decltype(&nvjpegDecode) Decode = nullptr;
    decltype(&nvjpegEncoderStateCreate) EncoderStateCreate = nullptr;
    decltype(&nvjpegEncoderParamsCreate) EncoderParamsCreate = nullptr;
    decltype(&nvjpegEncoderParamsSetQuality) SetQuality = nullptr;
    decltype(&nvjpegEncoderParamsSetSamplingFactors) SetSampling = nullptr;
    decltype(&nvjpegEncoderParamsSetOptimizedHuffman) SetOptimizedHuffman = nullptr;
    decltype(&nvjpegEncodeImage) EncodeImage = nullptr;
    decltype(&nvjpegEncodeRetrieveBitstream) RetrieveBitstream = nullptr;
    bool ok = false;
};

const Api& api() {
    static Api a = [] {
        Api x;
        for (const char* name : {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"}) {
            x.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (x.so) break;
        }
        if (!x.so) return x;
        auto sym = [&](const char* n) { return dlsym(x.so, n); };
#define VR180_SYM(field, name) x.field = reinterpret_cast<decltype(x.field)>(sym(#name))
        VR180_SYM(CreateSimple, nvjpegCreateSimple);
        VR180_SYM(JpegStateCreate, nvjpegJpegStateCreate);
        VR180_SYM(GetImageInfo, nvjpegGetImageInfo);
        VR180_SYM(Decode, nvjpegDecode);
        VR180_SYM(EncoderStateCreate, nvjpegEncoderStateCreate);
        VR180_SYM(EncoderParamsCreate, nvjpegEncoderParamsCreate);
        VR180_SYM(SetQuality, nvjpegEncoderParamsSetQuality);
        VR180_SYM(SetSampling, nvjpegEncoderParamsSetSamplingFactors);
        VR180_SYM(SetOptimizedHuffman, nvjpegEncoderParamsSetOptimizedHuffman);
        VR180_SYM(EncodeImage, nvjpegEncodeImage);
        VR180_SYM(RetrieveBitstream, nvjpegEncodeRetrieveBitstream);
#undef VR180_SYM
        x.ok = x.CreateSimple && x.JpegStateCreate && x.GetImageInfo && x.Decode && x.EncoderStateCreate &&
               x.EncoderParamsCreate && x.SetQuality && x.SetSampling && x.SetOptimizedHuffman && x.EncodeImage &&
               x.RetrieveBitstream;
        return x;
    }();
    return a;
}

// one nvJPEG handle + decoder / encoder state per host thread and device (the states are not thread safe)
struct ThreadState {
    int device = -1;
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t dec = nullptr;
    nvjpegEncoderState_t enc = nullptr;
    nvjpegEncoderParams_t params = nullptr;
};

int thread_state(ThreadState** out) {
    static thread_local std::vector<ThreadState> states;
    const Api& a = api();
    if (!a.ok) return VR180_ERR_UNSUPPORTED;
    int dev = 0;
    VR180_CUDA(cudaGetDevice(&dev));
    for (ThreadState& s : states)
        if (s.device == dev) {
            *out = &s;
            return VR180_OK;
        }
    ThreadState s;
    s.device = dev;
    if (a.CreateSimple(&s.handle) != NVJPEG_STATUS_SUCCESS) return VR180_ERR_UNSUPPORTED;
    if (a.JpegStateCreate(s.handle, &s.dec) != NVJPEG_STATUS_SUCCESS) return VR180_ERR_UNSUPPORTED;
    if (a.EncoderStateCreate(s.handle, &s.enc, nullptr) != NVJPEG_STATUS_SUCCESS) return VR180_ERR_UNSUPPORTED;
    if (a.EncoderParamsCreate(s.handle, &s.params, nullptr) != NVJPEG_STATUS_SUCCESS) return VR180_ERR_UNSUPPORTED;
    states.push_back(s);
    *out = &states.back();
    return VR180_OK;
}

}  // namespace

extern "C" {

int vr180_jpeg_available(void) { return api().ok ? 1 : 0; }

int vr180_jpeg_info(const uint8_t* jpeg_host, size_t n, int* width, int* height, int* channels) {
    if (!jpeg_host || n == 0 || !width || !height) return VR180_ERR_INVALID_ARG;
    ThreadState* ts = nullptr;
    const int rc = thread_state(&ts);
    if (rc != VR180_OK) return rc;
    int nc = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t css;
    if (api().GetImageInfo(ts->handle, jpeg_host, n, &nc, &css, w, h) != NVJPEG_STATUS_SUCCESS) return VR180_ERR_INVALID_ARG;
    *width = w[0];
    *height = h[0];
    if (channels) *channels = nc;
    return VR180_OK;
}

int vr180_jpeg_decode(const uint8_t* jpeg_host, size_t n, uint8_t* bgr_dev, int64_t pitch, int width, int height,
                      void* stream) {
    if (!jpeg_host || n == 0 || !bgr_dev || width <= 0 || height <= 0 || pitch < (int64_t)width * 3) return VR180_ERR_INVALID_ARG;
    DeviceGuard g(bgr_dev);
    if (!g.ok) return VR180_ERR_INVALID_ARG;
    ThreadState* ts = nullptr;
    const int rc = thread_state(&ts);
    if (rc != VR180_OK) return rc;
    int nc = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t css;
    if (api().GetImageInfo(ts->handle, jpeg_host, n, &nc, &css, w, h) != NVJPEG_STATUS_SUCCESS || w[0] != width || h[0] != height)
        return VR180_ERR_INVALID_ARG;
    nvjpegImage_t img;
    memset(&img, 0, sizeof(img));
    img.channel[0] = bgr_dev;
    img.pitch[0] = (size_t)pitch;
    // interleaved BGR, the layout cv.imread returns (grey-scale files are expanded to three channels, as cv.imread does)
    if (api().Decode(ts->handle, ts->dec, jpeg_host, n, NVJPEG_OUTPUT_BGRI, &img, (cudaStream_t)stream) != NVJPEG_STATUS_SUCCESS)
        return VR180_ERR_CUDA;
    return VR180_OK;
}

int vr180_jpeg_encode(const uint8_t* bgr_dev, int64_t pitch, int width, int height, int quality, uint8_t* out_host,
                      size_t* out_len, void* stream) {
    if (!bgr_dev || width <= 0 || height <= 0 || pitch < (int64_t)width * 3 || !out_len || quality < 1 || quality > 100)
        return VR180_ERR_INVALID_ARG;
    DeviceGuard g(bgr_dev);
    if (!g.ok) return VR180_ERR_INVALID_ARG;
    ThreadState* ts = nullptr;
    const int rc = thread_state(&ts);
    if (rc != VR180_OK) return rc;
    const Api& a = api();
    cudaStream_t st = (cudaStream_t)stream;
    // cv.imwrite's JPEG defaults: quality 95, 4:2:0 chroma, default (non-optimised) Huffman tables
    if (a.SetQuality(ts->params, quality, st) != NVJPEG_STATUS_SUCCESS ||
        a.SetSampling(ts->params, NVJPEG_CSS_420, st) != NVJPEG_STATUS_SUCCESS ||
        a.SetOptimizedHuffman(ts->params, 0, st) != NVJPEG_STATUS_SUCCESS)
        return VR180_ERR_CUDA;
    nvjpegImage_t img;
    memset(&img, 0, sizeof(img));
    img.channel[0] = const_cast<uint8_t*>(bgr_dev);
    img.pitch[0] = (size_t)pitch;
    if (a.EncodeImage(ts->handle, ts->enc, ts->params, &img, NVJPEG_INPUT_BGRI, width, height, st) != NVJPEG_STATUS_SUCCESS)
        return VR180_ERR_CUDA;
    size_t len = 0;
    if (a.RetrieveBitstream(ts->handle, ts->enc, nullptr, &len, st) != NVJPEG_STATUS_SUCCESS) return VR180_ERR_CUDA;
    VR180_CUDA(cudaStreamSynchronize(st));
    if (!out_host || *out_len < len) {  // report the size needed
        *out_len = len;
        return out_host ? VR180_ERR_NOMEM : VR180_OK;
    }
    if (a.RetrieveBitstream(ts->handle, ts->enc, out_host, &len, st) != NVJPEG_STATUS_SUCCESS) return VR180_ERR_CUDA;
    VR180_CUDA(cudaStreamSynchronize(st));
    *out_len = len;
    return VR180_OK;
}

}  // extern "C"
