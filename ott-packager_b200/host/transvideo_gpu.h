/*
 * transvideo_gpu.h -- host side (C) of the B200 transvideo hot path: ONE thread body that replaces the reference's
 * video_prepare_thread (source/transvideo.c:1896-2429) and video_scale_thread (1504-1683).  It keeps their contract:
 * frames arrive as dataqueue messages on preparevideo->input_queue, one message per rendition leaves on
 * encodevideo->input_queue[i], thumbnails leave on encodevideo->thumbnail_queue, buffers follow the mempool take/return
 * discipline, failures are reported through the same two hooks the reference uses (send_signal / send_direct_error).
 * All pixel work goes through libottgpu's C ABI (include/ottgpu.h); nothing here touches pixels.
 */
#ifndef OTTGPU_TRANSVIDEO_GPU_H_
#define OTTGPU_TRANSVIDEO_GPU_H_

#include "fillet_abi.h"
#include "../../include/ottgpu.h"

#if defined(__cplusplus)
extern "C" {
#endif

/* scale_struct of transvideo.c:124-127 (the thumbnail hand-off payload, consumed by video_thumbnail_thread 214-255) */
typedef struct ottgpu_scale_struct {
    uint8_t *output_data[4];
    int      output_stride[4];
} ottgpu_scale_struct;

typedef struct ottgpu_transvideo {
    /* ---- wiring: members of fillet_app_struct (include/fillet.h:368-454), filled by start_video_transcode_threads ---- */
    void *prepare_queue;                        /* core->preparevideo->input_queue                                   */
    void *encode_queue[OTTGPU_MAX_RUNGS];       /* core->encodevideo->input_queue[i]                                 */
    void *thumbnail_queue;                      /* core->encodevideo->thumbnail_queue                                */
    void *raw_video_pool;                       /* core->raw_video_pool                                              */
    void *msg_pool;                             /* core->fillet_msg_pool                                             */
    int   num_outputs;                          /* core->cd->num_outputs                                             */
    int   out_width[OTTGPU_MAX_RUNGS];          /* core->cd->transvideo_info[i].width                                */
    int   out_height[OTTGPU_MAX_RUNGS];         /* core->cd->transvideo_info[i].height                               */
    int   gpu;                                  /* core->cd->gpu                                                     */
    const int64_t *first_timestamp;             /* &vstream->first_timestamp (transvideo.c:2237-2241), may be NULL    */
    volatile int *running;                      /* video_prepare_thread_running                                      */
    /* ---- policy ---- */
    int   out_fmt;                              /* OTTGPU_FMT_I420 (x264/x265 consumers) | OTTGPU_FMT_NV12 (NVENC)    */
    int   device_outputs;                       /* 1: rendition buffers are device surfaces from `surface_pool[i]`
                                                      (msg->flags tagged OTTGPU_MSG_FLAG_DEVICE); 0: host, raw_video_pool */
    ottgpu_pool *surface_pool[OTTGPU_MAX_RUNGS];
    int   target;                               /* OTTGPU_TARGET_A = the reference's flags                           */
    /* ---- the reference's reporting hooks: send_signal / send_direct_error (include/esignal.h:79-80) ---- */
    void (*signal)(void *user, int signal_type, const char *message);
    void (*direct_error)(void *user, int signal_type, const char *message);   /* the reference follows it with exit(0) */
    void *user;                                 /* core                                                              */
    /* ---- state owned by the thread ---- */
    ottgpu_channel *channel;
    int   chan_width, chan_height, chan_deint;
    int64_t deinterlaced_frame_count;
    int   av_sync_compromised;
    int64_t frames_in, frames_out, thumbnails_out;
    /* ---- device-resident input frames (decode stage ran ottgpu_converter_convert with OTTGPU_MEM_DEVICE and tagged the
     *      message OTTGPU_MSG_FLAG_DEVICE): the pool they go back to once the deinterlacer window has moved past them ---- */
    ottgpu_pool *input_pool;
    void *held_inputs[8];
    int   n_held;
} ottgpu_transvideo;

/* pthread body: replaces video_prepare_thread + video_scale_thread.  context = ottgpu_transvideo*.
 * Returns NULL after draining the queue once *running becomes 0 (transvideo.c:1940-1952). */
void *ottgpu_video_prepare_scale_thread(void *context);

/* one message through the stage (what the loop above does per iteration); exposed for tests. 0 ok, <0 fatal */
int ottgpu_transvideo_process(ottgpu_transvideo *tv, dataqueue_message_struct *msg);

void ottgpu_transvideo_close(ottgpu_transvideo *tv);

#if defined(__cplusplus)
}
#endif
#endif
