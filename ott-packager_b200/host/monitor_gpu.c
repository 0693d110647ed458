/* See monitor_gpu.h.  Control flow follows source/transvideo.c:1685-1894 (video_monitor_thread); the per-frame
 * fprintf(stderr) of the reference is intentionally not reproduced. */
#include "monitor_gpu.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static void mon_fatal(ottgpu_monitor *mon, int code, const char *msg)
{
    if (mon->direct_error) mon->direct_error(mon->user, code, msg);
}

static void mon_free_saved(ottgpu_monitor *mon)
{
    if (!mon->saved_video_frame) return;
    if (mon->saved_is_device) ottgpu_pool_return(mon->device_pool, mon->saved_video_frame);
    else free(mon->saved_video_frame);
    mon->saved_video_frame = NULL;
}

int ottgpu_monitor_frame(ottgpu_monitor *mon, dataqueue_message_struct *msg)
{
    const int width = msg->width, height = msg->height;
    const int on_device = (msg->flags & OTTGPU_MSG_FLAG_DEVICE) == OTTGPU_MSG_FLAG_DEVICE;
    const size_t bytes = ((size_t)width * height * 3) / 2;

    mon->total_dead_frames = 0;                                           /* transvideo.c:1843 */
    if (!mon->monitor_ready || width != mon->monitor_width || height != mon->monitor_height ||
        on_device != mon->saved_is_device) {
        mon->monitor_width = width;
        mon->monitor_height = height;
        mon->monitor_num = msg->fps_num;
        mon->monitor_den = msg->fps_den;
        mon_free_saved(mon);
        mon->monitor_ready = 1;
    }
    if (!mon->saved_video_frame) {
        mon->saved_is_device = on_device;
        mon->saved_bytes = bytes;
        /* device streams keep the saved frame in a buffer of their own pool: filler frames then are plain D2D copies */
        mon->saved_video_frame = on_device ? ottgpu_pool_take(mon->device_pool) : malloc(bytes);
        if (!mon->saved_video_frame) {
            mon_fatal(mon, OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL, "Out of Uncompressed Video Buffers (MONITOR) - Restarting Service");
            return -1;
        }
    }
    if (on_device) {                                                      /* copy #2 (transvideo.c:1860) stays in HBM */
        if (ottgpu_memcpy(mon->gpu, mon->saved_video_frame, OTTGPU_MEM_DEVICE, msg->buffer, OTTGPU_MEM_DEVICE, bytes) != OTTGPU_OK) {
            mon_fatal(mon, OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL, ottgpu_last_error());
            return -1;
        }
    } else {
        memcpy(mon->saved_video_frame, msg->buffer, bytes);
    }
    mon->monitor_anchor_dts = msg->dts;
    mon->monitor_anchor_pts = msg->pts;
    mon->monitor_tff = msg->tff;
    mon->monitor_interlaced = msg->interlaced;
    mon->monitor_aspect_num = msg->aspect_num;
    mon->monitor_aspect_den = msg->aspect_den;
    /* the first real frame after filler frames makes the deinterlacer flush and reinitialise (transvideo.c:1875-1881) */
    msg->source_discontinuity = mon->dead_frames_triggered ? 1 : 0;
    mon->dead_frames_triggered = 0;
    dataqueue_put_front(mon->prepare_queue, msg);
    mon->frames_passed++;
    return 0;
}

int ottgpu_monitor_dead_time(ottgpu_monitor *mon)
{
    char fillermsg[256];
    double frame_delta, dead_frame_drift, dead_frames_actual;
    const double video_monitor_time = OTTGPU_VIDEO_MONITOR_DEAD_TIME;
    int64_t dead_frames, i;
    int dead_frame_adjustment;

    if (!mon->monitor_ready || !mon->saved_video_frame || mon->monitor_num <= 0 || mon->monitor_den <= 0) return 0;
    mon->total_video_monitor_dead_time += (int64_t)video_monitor_time;
    frame_delta = 90000.0 / ((double)mon->monitor_num / (double)mon->monitor_den);
    dead_frame_drift = mon->expected_dead_frames - (double)mon->total_dead_frames;
    dead_frame_adjustment = dead_frame_drift >= 1 ? 1 : 0;
    dead_frames_actual = video_monitor_time / (1000000.0 / ((double)mon->monitor_num / (double)mon->monitor_den));
    mon->expected_dead_frames += dead_frames_actual;
    dead_frames = (int64_t)dead_frames_actual + dead_frame_adjustment;
    mon->dead_frames_triggered = 1;
    if (mon->reinitialize_decoder) *mon->reinitialize_decoder = 1;
    snprintf(fillermsg, sizeof fillermsg - 1, "Signal Loss Inserting %ld Filler Video Frames, Frame Delta %f, Framerate %d/%d",
             (long)dead_frames, frame_delta, mon->monitor_num, mon->monitor_den);
    if (mon->signal) mon->signal(mon->user, OTTGPU_SIGNAL_FRAME_VIDEO_FILLER, fillermsg);

    for (i = 0; i < dead_frames; i++) {
        dataqueue_message_struct *prepare_msg;
        void *output_video_frame;
        mon->total_dead_frames++;
        output_video_frame = mon->saved_is_device ? ottgpu_pool_take(mon->device_pool)
                                                  : memory_take(mon->raw_video_pool, (int)mon->saved_bytes);
        if (!output_video_frame) {
            mon_fatal(mon, OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL, "Out of Uncompressed Video Buffers (MONITOR) - Restarting Service");
            return -1;
        }
        prepare_msg = (dataqueue_message_struct *)memory_take(mon->msg_pool, sizeof(dataqueue_message_struct));
        if (!prepare_msg) {
            if (mon->saved_is_device) ottgpu_pool_return(mon->device_pool, output_video_frame);
            else memory_return(mon->raw_video_pool, output_video_frame);
            mon_fatal(mon, OTTGPU_SIGNAL_DIRECT_ERROR_MSGPOOL, "Out of Message Video Buffers (MONITOR) - Restarting Service");
            return -1;
        }
        memset(prepare_msg, 0, sizeof(*prepare_msg));
        if (mon->saved_is_device) {                                       /* transvideo.c:1774, in HBM */
            ottgpu_memcpy(mon->gpu, output_video_frame, OTTGPU_MEM_DEVICE, mon->saved_video_frame, OTTGPU_MEM_DEVICE, mon->saved_bytes);
            prepare_msg->flags = OTTGPU_MSG_FLAG_DEVICE | ((uint32_t)mon->gpu << 8);
        } else {
            memcpy(output_video_frame, mon->saved_video_frame, mon->saved_bytes);
        }
        prepare_msg->buffer = output_video_frame;
        prepare_msg->buffer_size = mon->saved_bytes;
        prepare_msg->pts = (int64_t)((double)mon->monitor_anchor_pts + ((double)mon->total_dead_frames * frame_delta));
        prepare_msg->dts = (int64_t)((double)mon->monitor_anchor_dts + ((double)mon->total_dead_frames * frame_delta));
        prepare_msg->tff = (uint8_t)mon->monitor_tff;
        prepare_msg->interlaced = (uint8_t)mon->monitor_interlaced;
        prepare_msg->fps_num = mon->monitor_num;
        prepare_msg->fps_den = mon->monitor_den;
        prepare_msg->aspect_num = mon->monitor_aspect_num;
        prepare_msg->aspect_den = mon->monitor_aspect_den;
        prepare_msg->width = mon->monitor_width;
        prepare_msg->height = mon->monitor_height;
        prepare_msg->source_discontinuity = 0;
        dataqueue_put_front(mon->prepare_queue, prepare_msg);
        mon->filler_frames++;
    }
    return (int)dead_frames;
}

static double mon_elapsed_us(const struct timespec *a, const struct timespec *b)
{
    return (double)(b->tv_sec - a->tv_sec) * 1e6 + (double)(b->tv_nsec - a->tv_nsec) / 1e3;
}

void *ottgpu_video_monitor_thread(void *context)
{
    ottgpu_monitor *mon = (ottgpu_monitor *)context;
    struct timespec monitor_start, monitor_end;
    if (mon->reinitialize_decoder) *mon->reinitialize_decoder = 0;
    while (*mon->running) {
        dataqueue_message_struct *msg;
        clock_gettime(CLOCK_MONOTONIC, &monitor_start);
        msg = dataqueue_take_back(mon->monitor_queue);
        while (!msg && *mon->running) {
            usleep(1000);
            if (mon->monitor_ready) {
                clock_gettime(CLOCK_MONOTONIC, &monitor_end);
                if (mon_elapsed_us(&monitor_start, &monitor_end) >= OTTGPU_VIDEO_MONITOR_DEAD_TIME) {
                    clock_gettime(CLOCK_MONOTONIC, &monitor_start);
                    if (ottgpu_monitor_dead_time(mon) < 0 && !*mon->running) break;
                }
            }
            msg = dataqueue_take_back(mon->monitor_queue);
        }
        if (!*mon->running) {                                             /* drain, transvideo.c:1821-1836 */
            while (msg) {
                free(msg->caption_buffer);
                msg->caption_buffer = NULL;
                if ((msg->flags & OTTGPU_MSG_FLAG_DEVICE) == OTTGPU_MSG_FLAG_DEVICE) ottgpu_pool_return(mon->device_pool, msg->buffer);
                else memory_return(mon->raw_video_pool, msg->buffer);
                msg->buffer = NULL;
                memory_return(mon->msg_pool, msg);
                msg = dataqueue_take_back(mon->monitor_queue);
            }
            break;
        }
        if (msg) ottgpu_monitor_frame(mon, msg);
    }
    ottgpu_monitor_close(mon);
    return NULL;
}

void ottgpu_monitor_close(ottgpu_monitor *mon)
{
    mon_free_saved(mon);
}
