// Device orchestration of the JPEG block stage (see jpeg.h).
#include <cmath>
#include <cstring>

#include "jpeg.h"

namespace uhdr_b200 {

int jpeg_forward_dev(Workspace& ws, const DevImage& img, int quality, JpegEncodeJob* job, bool zigzag) {
  int rc = jpeg_frame_init(&job->frame, img.v.fmt, img.v.w, img.v.h, quality);
  if (rc) return rc;
  const JpegFrame& f = job->frame;
  job->zigzag = zigzag;
  for (int c = 0; c < f.ncomp; c++) {
    const JpegComp& k = f.comp[c];
    job->d_coefs[c] = (int16_t*)ws.dalloc(f.blocks(c) * 128);
    if (!job->d_coefs[c]) return E_MEM;
    DctPlaneParams p;
    memset(&p, 0, sizeof p);
    p.wblocks = k.wblocks;
    p.hblocks = k.hblocks;
    p.coefs = job->d_coefs[c];
    p.zigzag_out = zigzag ? 1 : 0;
    memcpy(p.q, f.qt[k.tq], sizeof p.q);
    if (img.v.fmt == F_RGB888) {
      // jpeg_write_scanlines path: jccolor.c conversion, edges replicated (jcsample.c/jcprepct.c)
      p.src = (const uint8_t*)img.v.p[0];
      p.src_stride = img.v.stride[0];
      p.w = img.v.w;
      p.h = img.v.h;
      p.pad_mode = 1;
      p.rgb_comp = c;
    } else {
      // raw_data_in path (jpegencoderhelper.cpp:246-309): whole blocks are read from the plane
      // (device strides are >= wblocks*8 and the bytes past the width are defined, see
      // alloc_dev_image / upload); rows past the plane height come from the helper's pad row:
      // 0 for luma, 128 for chroma.
      if (img.v.stride[c] < k.wblocks * 8)
        return fail(E_ERROR, "internal: device plane stride %d < padded width %d", img.v.stride[c], k.wblocks * 8);
      p.src = (const uint8_t*)img.v.p[c];
      p.src_stride = img.v.stride[c];
      p.w = k.wblocks * 8;
      p.h = k.height;
      p.pad_mode = 0;
      p.fill = c == 0 ? 0 : 128;
      p.rgb_comp = -1;
    }
    TIMED(ws, "fdct_quant", launch_fdct_quant(p, ws.stream()));
  }
  return E_OK;
}

int jpeg_fetch_coefs(Workspace& ws, JpegEncodeJob* job) {
  const JpegFrame& f = job->frame;
  for (int c = 0; c < f.ncomp; c++) {
    job->h_coefs[c] = (int16_t*)ws.halloc(f.blocks(c) * 128);
    if (!job->h_coefs[c]) return E_MEM;
    CUDA_TRY(cudaMemcpyAsync(job->h_coefs[c], job->d_coefs[c], f.blocks(c) * 128, cudaMemcpyDeviceToHost, ws.stream()));
  }
  return E_OK;
}

int jpeg_inverse_dev(Workspace& ws, const JpegHeader& h, int16_t* const h_coefs[3], uint8_t* d_planes[3],
                     int plane_stride[3]) {
  const JpegFrame& f = h.frame;
  for (int c = 0; c < f.ncomp; c++) {
    const JpegComp& k = f.comp[c];
    int16_t* d = (int16_t*)ws.dalloc(f.blocks(c) * 128);
    if (!d) return E_MEM;
    CUDA_TRY(cudaMemcpyAsync(d, h_coefs[c], f.blocks(c) * 128, cudaMemcpyHostToDevice, ws.stream()));
    IdctPlaneParams p;
    memset(&p, 0, sizeof p);
    p.coefs = d;
    memcpy(p.q, f.qt[k.tq], sizeof p.q);
    p.wblocks = k.wblocks;
    p.hblocks = k.hblocks;
    p.dst = d_planes[c];
    p.dst_stride = plane_stride[c];
    p.dst_w = k.wblocks * 8 < plane_stride[c] ? k.wblocks * 8 : plane_stride[c];
    p.dst_h = k.hblocks * 8;
    TIMED(ws, "idct_dequant", launch_idct_dequant(p, ws.stream()));
  }
  return E_OK;
}

}  // namespace uhdr_b200
