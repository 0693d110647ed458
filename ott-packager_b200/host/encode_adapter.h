/*
 * encode_adapter.h -- the encoder threads' side of the rung hand-off: what video_encode_thread_x264 /
 * _x265 / _nvenc do with one rendition message once the rungs are produced by libottgpu.
 *
 *   x264  (source/transvideo.c:1330-1338): three memcpy()s (copy #6) from the contiguous pool buffer into pic.img.plane[].
 *   x265  (960-967): plane pointers straight into the contiguous buffer, strides w, w/2, w/2.
 *   NVENC (532-557, 587): CPU NV12 interleave into encode_av_frame + av_hwframe_transfer_data (H2D) per rendition.
 *
 * With libottgpu the message either carries a HOST buffer in exactly the reference's contiguous w*h*3/2 layout (the
 * library copied the rung into a pinned pool buffer - nothing changes for x264/x265 except that copy #5 is gone), or a
 * DEVICE surface (msg->flags & OTTGPU_MSG_FLAG_DEVICE).  A CPU encoder maps a device surface through a pinned staging
 * buffer (one D2H); an NVENC session registers the surface itself: the interleave and the H2D transfer disappear.
 */
#ifndef OTTGPU_ENCODE_ADAPTER_H_
#define OTTGPU_ENCODE_ADAPTER_H_

#include "fillet_abi.h"
#include "../../include/ottgpu.h"

#if defined(__cplusplus)
extern "C" {
#endif

typedef struct ottgpu_encode_adapter {
    /* ---- wiring ---- */
    void *raw_video_pool;                       /* core->raw_video_pool: where host rendition buffers go back to       */
    void *msg_pool;                             /* core->fillet_msg_pool                                             */
    ottgpu_pool *surface_pool;                  /* device surfaces of THIS rendition (the scale stage took them here)  */
    int   gpu;
    int   width, height;                        /* core->cd->transvideo_info[i].width / height                        */
    int   fmt;                                  /* OTTGPU_FMT_I420 | NV12 | P010 | I420P10 of the rendition buffers    */
    int   pitch_align;                          /* as in the rung configuration                                      */
    /* ---- state ---- */
    ottgpu_pool *staging_pool;                  /* one pinned buffer                                                  */
    void *staging;                              /* pinned host copy of a device surface for a CPU encoder             */
    size_t staging_bytes;
    int64_t frames, d2h_frames;
} ottgpu_encode_adapter;

/* picture description in the encoder's terms: x264_picture_t.img.plane / i_stride, x265_picture.planes / stride */
typedef struct ottgpu_encode_planes {
    const uint8_t *plane[3];
    int stride[3];
    int planes;
} ottgpu_encode_planes;

/* x264 / x265: plane pointers + strides of the rendition in HOST memory.  Host messages: pointers into msg->buffer
 * (zero copy, exactly transvideo.c:960-967).  Device messages: one D2H of the surface into the adapter's pinned staging
 * buffer, valid until the next call. */
int ottgpu_encode_map_host(ottgpu_encode_adapter *ad, const dataqueue_message_struct *msg, ottgpu_encode_planes *out);

/* NVENC: what cuda-based encoder input registration needs for a device surface (NV_ENC_REGISTER_RESOURCE with
 * resourceType CUDADEVICEPTR: device pointer, pitch, width, height, buffer format).  Fails for host messages. */
typedef struct ottgpu_nvenc_surface {
    void *device_ptr;                           /* start of the luma plane                                            */
    size_t chroma_offset;                       /* byte offset of the interleaved UV plane                            */
    int   pitch;                                /* bytes, both planes                                                 */
    int   width, height;
    int   fmt;                                  /* OTTGPU_FMT_NV12 | OTTGPU_FMT_P010                                  */
    int   gpu;
} ottgpu_nvenc_surface;
int ottgpu_encode_nvenc_surface(const ottgpu_encode_adapter *ad, const dataqueue_message_struct *msg, ottgpu_nvenc_surface *out);

/* the end of every encoder thread's loop body (transvideo.c:1456-1461, 1081-1086, 715-720): buffer back to where it came
 * from, message back to the message pool */
void ottgpu_encode_release(ottgpu_encode_adapter *ad, dataqueue_message_struct *msg);

/* is there an NVENC on this box?  dlopen("libnvidia-encode.so.1") + NvEncodeAPIGetMaxSupportedVersion; writes a one-line
 * report into buf, returns 1 if an encoder API is present, 0 if not */
int ottgpu_nvenc_probe(char *buf, size_t size);

void ottgpu_encode_adapter_close(ottgpu_encode_adapter *ad);

#if defined(__cplusplus)
}
#endif
#endif
