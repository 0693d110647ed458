/* thumbnail_gpu.c -- see thumbnail_gpu.h.  Mirrors source/transvideo.c:200-262 (thread loop) and 164-198 (the encode). */
#include "thumbnail_gpu.h"
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

int ottgpu_thumbnail_process(ottgpu_thumbnailer *th, dataqueue_message_struct *msg)
{
    int rc = 0;
    ottgpu_scale_struct *thumbnail_output = (ottgpu_scale_struct *)msg->buffer;     /* transvideo.c:211 */
    if (thumbnail_output) {
        if (!th->encoder) {
            rc = ottgpu_jpeg_create(th->gpu, OTTGPU_THUMBNAIL_WIDTH, OTTGPU_THUMBNAIL_HEIGHT, th->qscale > 0 ? th->qscale : 3, &th->encoder);
            if (rc == 0) {
                th->file_capacity = ottgpu_jpeg_max_bytes(th->encoder);
                th->file_buffer = malloc(th->file_capacity);
                if (!th->file_buffer) rc = OTTGPU_ERR_NOMEM;
            }
        }
        if (rc == 0) {
            ottgpu_picture pic;                                                     /* jpeg_frame->data / linesize, 218-225 */
            for (int p = 0; p < 3; ++p) {
                pic.plane[p] = thumbnail_output->output_data[p];
                pic.stride[p] = thumbnail_output->output_stride[p];
            }
            pic.mem = OTTGPU_MEM_HOST;
            size_t bytes = 0;
            rc = ottgpu_jpeg_encode(th->encoder, &pic, th->file_buffer, th->file_capacity, &bytes);
            if (rc == 0) {
                char filename[512];
                snprintf(filename, sizeof(filename) - 1, "%s/thumbnail%d.jpg", th->directory ? th->directory : "/var/www/html", th->identity);
                FILE *JPEG = fopen(filename, "wb");                                 /* 188-192: a failed open is not an error there either */
                if (JPEG) {
                    fwrite(th->file_buffer, 1, bytes, JPEG);
                    fclose(JPEG);
                }
                th->last_bytes = bytes;
                th->written++;
            }
        }
        if (rc != 0) th->failed++;
        free(thumbnail_output->output_data[0]);                                     /* av_freep(&output_data[0]), 254 */
        free(thumbnail_output);                                                     /* 255 */
    }
    memory_return(th->msg_pool, msg);                                               /* 258 */
    return rc;
}

void *ottgpu_video_thumbnail_thread(void *context)
{
    ottgpu_thumbnailer *th = (ottgpu_thumbnailer *)context;
    while (*th->running) {
        dataqueue_message_struct *msg = dataqueue_take_back(th->thumbnail_queue);
        while (!msg && *th->running) {
            usleep(10000);                                                          /* 207 */
            msg = dataqueue_take_back(th->thumbnail_queue);
        }
        if (!msg) break;
        ottgpu_thumbnail_process(th, msg);
    }
    return NULL;
}

void ottgpu_thumbnailer_close(ottgpu_thumbnailer *th)
{
    if (!th) return;
    if (th->encoder) ottgpu_jpeg_destroy(th->encoder);
    th->encoder = NULL;
    free(th->file_buffer);
    th->file_buffer = NULL;
}
