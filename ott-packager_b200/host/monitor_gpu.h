/*
 * monitor_gpu.h -- the signal-loss filler of video_monitor_thread (source/transvideo.c:1685-1894) with the saved frame
 * and the filler frames kept where the stream's frames live.  The reference memcpy()s every passing frame into
 * saved_video_frame (copy #2, transvideo.c:1860) and, after one second without input, memcpy()s it again into one fresh
 * pool buffer per filler frame (1774).  Here a device-resident stream (messages tagged OTTGPU_MSG_FLAG_DEVICE by the
 * decode-side converter) keeps its saved frame in HBM: both copies become device-to-device copies and nothing crosses
 * PCIe.  Host frames behave exactly like the reference.  Arithmetic (frame counts, drift correction, pts/dts anchors,
 * discontinuity flag on the first real frame after a gap) follows the reference line by line.
 */
#ifndef OTTGPU_MONITOR_GPU_H_
#define OTTGPU_MONITOR_GPU_H_

#include "fillet_abi.h"
#include "../../include/ottgpu.h"

#if defined(__cplusplus)
extern "C" {
#endif

#define OTTGPU_SIGNAL_FRAME_VIDEO_FILLER 0x18          /* esignal.h:53 */
#define OTTGPU_VIDEO_MONITOR_DEAD_TIME   1000000       /* microseconds, transvideo.c:1730 */

typedef struct ottgpu_monitor {
    /* ---- wiring (members of fillet_app_struct) ---- */
    void *monitor_queue;                        /* core->monitorvideo->input_queue                                   */
    void *prepare_queue;                        /* core->preparevideo->input_queue                                   */
    void *raw_video_pool;                       /* core->raw_video_pool                                              */
    void *msg_pool;                             /* core->fillet_msg_pool                                             */
    ottgpu_pool *device_pool;                   /* where device-resident filler frames come from (the converter's pool) */
    int   gpu;
    volatile int *running;                      /* video_monitor_thread_running                                      */
    volatile int *reinitialize_decoder;         /* &core->reinitialize_decoder (transvideo.c:1754), may be NULL        */
    void (*signal)(void *user, int signal_type, const char *message);
    void (*direct_error)(void *user, int signal_type, const char *message);
    void *user;
    /* ---- state (the locals of video_monitor_thread) ---- */
    int   monitor_ready, monitor_width, monitor_height, monitor_num, monitor_den;
    int   monitor_tff, monitor_interlaced, monitor_aspect_num, monitor_aspect_den;
    int64_t monitor_anchor_dts, monitor_anchor_pts;
    int64_t total_dead_frames;
    int64_t total_video_monitor_dead_time;
    double expected_dead_frames;
    int   dead_frames_triggered;
    void *saved_video_frame;                    /* host malloc() or device memory, see saved_is_device                */
    int   saved_is_device;
    size_t saved_bytes;
    int64_t frames_passed, filler_frames;
} ottgpu_monitor;

/* one real frame through the stage (transvideo.c:1839-1884): saves it, stamps source_discontinuity, forwards it */
int ottgpu_monitor_frame(ottgpu_monitor *mon, dataqueue_message_struct *msg);
/* the dead-time branch (transvideo.c:1726-1812), entered when no frame arrived for OTTGPU_VIDEO_MONITOR_DEAD_TIME us;
 * returns the number of filler frames queued, <0 after a fatal report */
int ottgpu_monitor_dead_time(ottgpu_monitor *mon);
/* pthread body replacing video_monitor_thread; context = ottgpu_monitor* */
void *ottgpu_video_monitor_thread(void *context);
void ottgpu_monitor_close(ottgpu_monitor *mon);

#if defined(__cplusplus)
}
#endif
#endif
