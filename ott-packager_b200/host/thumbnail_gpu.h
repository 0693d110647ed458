/*
 * thumbnail_gpu.h -- video_thumbnail_thread (source/transvideo.c:200-262) with save_frame_as_jpeg (164-198) on the GPU
 * (SURVEY 8(f) N4).  The reference wraps the scale_struct of the thumbnail hand-off in an AVFrame, opens a fresh libavcodec
 * MJPEG context per picture, encodes one YUVJ420P frame and writes /var/www/html/thumbnail<identity>.jpg; then it frees the
 * picture (av_freep(&output_data[0]); free(struct)) and returns the message to fillet_msg_pool.  Here the encode is
 * ottgpu_jpeg_encode (one encoder kept for the life of the thread); queue, ownership and file conventions are unchanged.
 */
#ifndef OTTGPU_THUMBNAIL_GPU_H_
#define OTTGPU_THUMBNAIL_GPU_H_

#include "fillet_abi.h"
#include "transvideo_gpu.h"
#include "../../include/ottgpu.h"

#if defined(__cplusplus)
extern "C" {
#endif

#define OTTGPU_THUMBNAIL_WIDTH  176                 /* THUMBNAIL_WIDTH / THUMBNAIL_HEIGHT, transvideo.c:71-72 */
#define OTTGPU_THUMBNAIL_HEIGHT 144

typedef struct ottgpu_thumbnailer {
    /* ---- wiring ---- */
    void *thumbnail_queue;                      /* core->encodevideo->thumbnail_queue                                */
    void *msg_pool;                             /* core->fillet_msg_pool                                             */
    int   gpu;
    int   identity;                             /* core->cd->identity: the file is thumbnail<identity>.jpg            */
    const char *directory;                      /* NULL = "/var/www/html" (transvideo.c:187)                          */
    int   qscale;                               /* 0 = 3: what libavcodec's rate control picks for natural content at the
                                                   reference's 250 kb/s (transvideo.c:175)                            */
    volatile int *running;                      /* video_thumbnail_thread_running                                    */
    /* ---- state ---- */
    ottgpu_jpeg *encoder;
    void *file_buffer;
    size_t file_capacity;
    int64_t written, failed;
    size_t last_bytes;
} ottgpu_thumbnailer;

/* one message off the thumbnail queue: encode, write the file, free the picture, return the message (transvideo.c:211-258) */
int ottgpu_thumbnail_process(ottgpu_thumbnailer *th, dataqueue_message_struct *msg);
/* pthread body replacing video_thumbnail_thread; context = ottgpu_thumbnailer* */
void *ottgpu_video_thumbnail_thread(void *context);
void ottgpu_thumbnailer_close(ottgpu_thumbnailer *th);

#if defined(__cplusplus)
}
#endif
#endif
