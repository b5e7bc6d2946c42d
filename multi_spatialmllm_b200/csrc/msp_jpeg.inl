// JPEG decode / encode of BGR frames for the `*_dot` streams — thin glue over nvJPEG (library code, not a kernel of
// ours): the reference does `cv2.imread` -> annotate -> `cv2.imwrite` per sample (DE-dot:221-229, DC-dot:333-346,
// VC-dot:357-397); here frames are decoded to HBM, painted by k_draw_marks and encoded without a host round trip.
// libnvjpeg is opened lazily with dlopen so that the rest of libmsp.so never depends on it.
// Included at the end of msp_kernels.cu (shares its error plumbing).
#include <dlfcn.h>
#include <nvjpeg.h>

namespace {

struct NvJpegApi {
  decltype(&nvjpegCreateSimple) CreateSimple;
  decltype(&nvjpegJpegStateCreate) JpegStateCreate;
  decltype(&nvjpegGetImageInfo) GetImageInfo;
  decltype(&nvjpegDecode) Decode;
  decltype(&nvjpegEncoderStateCreate) EncoderStateCreate;
  decltype(&nvjpegEncoderParamsCreate) EncoderParamsCreate;
  decltype(&nvjpegEncoderParamsSetQuality) EncoderParamsSetQuality;
  decltype(&nvjpegEncoderParamsSetSamplingFactors) EncoderParamsSetSamplingFactors;
  decltype(&nvjpegEncoderParamsSetOptimizedHuffman) EncoderParamsSetOptimizedHuffman;
  decltype(&nvjpegEncodeImage) EncodeImage;
  decltype(&nvjpegEncodeRetrieveBitstream) EncodeRetrieveBitstream;
  bool ok = false;
};

const NvJpegApi* nvjpeg_api() {
  static NvJpegApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;   // benign race: every thread resolves the same pointers
  void* so = nullptr;
  const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
  for (const char* n : names)
    if ((so = dlopen(n, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
  bool ok = so != nullptr;
#define MSP_NVJ(sym) ok = ok && ((api.sym = reinterpret_cast<decltype(api.sym)>(dlsym(so, "nvjpeg" #sym))) != nullptr)
  MSP_NVJ(CreateSimple);
  MSP_NVJ(JpegStateCreate);
  MSP_NVJ(GetImageInfo);
  MSP_NVJ(Decode);
  MSP_NVJ(EncoderStateCreate);
  MSP_NVJ(EncoderParamsCreate);
  MSP_NVJ(EncoderParamsSetQuality);
  MSP_NVJ(EncoderParamsSetSamplingFactors);
  MSP_NVJ(EncoderParamsSetOptimizedHuffman);
  MSP_NVJ(EncodeImage);
  MSP_NVJ(EncodeRetrieveBitstream);
#undef MSP_NVJ
  api.ok = ok;
  tried = true;
  return ok ? &api : nullptr;
}

struct JpegCtx {  // nvJPEG states are not thread safe: one set per host thread
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t dec = nullptr;
  nvjpegEncoderState_t enc = nullptr;
  nvjpegEncoderParams_t params = nullptr;
};
thread_local JpegCtx t_jpeg;

int jpeg_fail(const char* what, int status) {
  snprintf(g_err, sizeof(g_err), "%s: nvjpeg status %d", what, status);
  return MSP_ELAUNCH;
}

int jpeg_ctx(const NvJpegApi*& api, cudaStream_t s) {
  api = nvjpeg_api();
  if (api == nullptr) return fail(MSP_EINVAL, "libnvjpeg.so.12 could not be opened (JPEG entry points unavailable)");
  JpegCtx& c = t_jpeg;
  nvjpegStatus_t st;
  if (c.handle == nullptr && (st = api->CreateSimple(&c.handle)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegCreateSimple", st);
  if (c.dec == nullptr && (st = api->JpegStateCreate(c.handle, &c.dec)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegJpegStateCreate", st);
  if (c.enc == nullptr && (st = api->EncoderStateCreate(c.handle, &c.enc, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegEncoderStateCreate", st);
  if (c.params == nullptr && (st = api->EncoderParamsCreate(c.handle, &c.params, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegEncoderParamsCreate", st);
  return MSP_OK;
}

}  // namespace

extern "C" int msp_jpeg_info(const uint8_t* host_jpeg, size_t length, int32_t* height, int32_t* width) {
  REQUIRE(host_jpeg && length > 0 && height && width, "msp_jpeg_info: NULL pointer or empty stream");
  const NvJpegApi* api;
  if (int rc = jpeg_ctx(api, nullptr)) return rc;
  int comps = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t css;
  const nvjpegStatus_t st = api->GetImageInfo(t_jpeg.handle, host_jpeg, length, &comps, &css, ws, hs);
  if (st != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegGetImageInfo", st);
  *height = hs[0];
  *width = ws[0];
  return MSP_OK;
}

extern "C" int msp_jpeg_decode_bgr(const uint8_t* host_jpeg, size_t length, uint8_t* image_out, int32_t H, int32_t W,
                                   void* stream) {
  REQUIRE(host_jpeg && length > 0 && image_out && H > 0 && W > 0, "msp_jpeg_decode_bgr: bad argument");
  const NvJpegApi* api;
  if (int rc = jpeg_ctx(api, (cudaStream_t)stream)) return rc;
  int32_t h = 0, w = 0;
  if (int rc = msp_jpeg_info(host_jpeg, length, &h, &w)) return rc;
  REQUIRE(h == H && w == W, "msp_jpeg_decode_bgr: the stream's size differs from (H, W)");
  nvjpegImage_t dst;
  memset(&dst, 0, sizeof(dst));
  dst.channel[0] = image_out;
  dst.pitch[0] = (size_t)W * 3;
  const nvjpegStatus_t st = api->Decode(t_jpeg.handle, t_jpeg.dec, host_jpeg, length, NVJPEG_OUTPUT_BGRI, &dst, (cudaStream_t)stream);
  if (st != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegDecode", st);
  return MSP_OK;
}

extern "C" int msp_jpeg_encode_bgr(const uint8_t* image, int32_t H, int32_t W, int32_t quality, uint8_t* host_out,
                                   size_t capacity, size_t* length, void* stream) {
  REQUIRE(image && H > 0 && W > 0 && length, "msp_jpeg_encode_bgr: bad argument");
  REQUIRE(quality >= 1 && quality <= 100, "msp_jpeg_encode_bgr: quality must be in [1, 100]");
  const NvJpegApi* api;
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = jpeg_ctx(api, s)) return rc;
  JpegCtx& c = t_jpeg;
  nvjpegStatus_t st;
  if ((st = api->EncoderParamsSetQuality(c.params, quality, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("SetQuality", st);
  if ((st = api->EncoderParamsSetSamplingFactors(c.params, NVJPEG_CSS_420, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("SetSamplingFactors", st);
  if ((st = api->EncoderParamsSetOptimizedHuffman(c.params, 0, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("SetOptimizedHuffman", st);
  nvjpegImage_t src;
  memset(&src, 0, sizeof(src));
  src.channel[0] = const_cast<uint8_t*>(image);
  src.pitch[0] = (size_t)W * 3;
  if ((st = api->EncodeImage(c.handle, c.enc, c.params, &src, NVJPEG_INPUT_BGRI, W, H, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegEncodeImage", st);
  size_t need = 0;
  if ((st = api->EncodeRetrieveBitstream(c.handle, c.enc, nullptr, &need, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegEncodeRetrieveBitstream(size)", st);
  *length = need;
  if (host_out == nullptr) return MSP_OK;                       // size query
  REQUIRE(need <= capacity, "msp_jpeg_encode_bgr: output buffer too small (*length holds the size needed)");
  if ((st = api->EncodeRetrieveBitstream(c.handle, c.enc, host_out, &need, s)) != NVJPEG_STATUS_SUCCESS) return jpeg_fail("nvjpegEncodeRetrieveBitstream", st);
  if (cudaStreamSynchronize(s) != cudaSuccess) return check_launch("msp_jpeg_encode_bgr");
  *length = need;
  return MSP_OK;
}
