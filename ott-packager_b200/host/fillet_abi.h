/*
 * fillet_abi.h -- the slice of ott-packager's C interfaces the transvideo hot path talks to, declared here so that
 * transvideo_gpu.c builds both inside the reference tree and stand-alone (tests).  Nothing is implemented here:
 *   - dataqueue_message_struct mirrors include/dataqueue.h:32-59 field for field (400 bytes on LP64); the static
 *     assertions below and tests/test_host_glue.py (which compiles against the real header when /root/reference is
 *     present) keep it honest.  When OTTGPU_USE_FILLET_HEADERS is defined the reference's own headers are used instead.
 *   - dataqueue_* / memory_* are the reference's functions (source/dataqueue.c, source/mempool.c), linked unchanged.
 */
#ifndef OTTGPU_FILLET_ABI_H_
#define OTTGPU_FILLET_ABI_H_

#if defined(OTTGPU_USE_FILLET_HEADERS)
#include "dataqueue.h"
#include "mempool.h"
#else
#include <stddef.h>
#include <stdint.h>

#define MAX_SMALLBUF_SIZE 256

/* same members, order and types as include/dataqueue.h:32-59; byte offsets on LP64 in the right-hand column */
typedef struct _dataqueue_message_struct_ {
    int buffer_type;  unsigned long buffer_size;                                         /*   0    8          */
    int64_t pts, dts;                                                                    /*  16   24          */
    uint32_t flags;  int source_discontinuity;  uint8_t tff, interlaced;                 /*  32   36   40  41 */
    int fps_num, fps_den, aspect_num, aspect_den, width, height;                         /*  44 .. 64         */
    uint8_t stream_index;  int channels, sample_rate;                                    /*  68   72   76     */
    int64_t first_pts;  int caption_size;  int64_t caption_timestamp;                    /*  80   88   96     */
    int splice_point;  int64_t splice_duration, splice_duration_remaining;               /* 104  112  120     */
    char smallbuf[MAX_SMALLBUF_SIZE];  uint8_t *caption_buffer;  void *buffer;           /* 128  384  392     */
} dataqueue_message_struct;

#if defined(__LP64__) && defined(__STDC_VERSION__) && __STDC_VERSION__ >= 201112L
_Static_assert(sizeof(dataqueue_message_struct) == 400, "dataqueue_message_struct must stay 400 bytes (include/dataqueue.h)");
_Static_assert(offsetof(dataqueue_message_struct, flags) == 32 && offsetof(dataqueue_message_struct, width) == 60 &&
               offsetof(dataqueue_message_struct, smallbuf) == 128 && offsetof(dataqueue_message_struct, buffer) == 392,
               "dataqueue_message_struct layout drifted from include/dataqueue.h");
#endif

#if defined(__cplusplus)
extern "C" {
#endif
/* include/dataqueue.h:65-69 */
void *dataqueue_create(void);
int dataqueue_destroy(void *queue);
int dataqueue_get_size(void *queue);
int dataqueue_put_front(void *queue, dataqueue_message_struct *message);
dataqueue_message_struct *dataqueue_take_back(void *queue);
/* include/mempool.h:30-35 */
void *memory_create(int buffer_count, int buffer_size);
int memory_destroy(void *pool);
void *memory_take(void *pool, int owner);
int memory_return(void *pool, void *buffer);
int memory_reset(void *pool);
int memory_unused(void *pool);
#if defined(__cplusplus)
}
#endif
#endif /* OTTGPU_USE_FILLET_HEADERS */

/* "device resident" tag for dataqueue_message_struct.flags (unused by the video path today, include/dataqueue.h:37):
 * msg->buffer is a device pointer of the GPU in bits 8..15 and must be returned to an ottgpu_pool, not to mempool. */
#define OTTGPU_MSG_FLAG_DEVICE   0x4f540001u
#define OTTGPU_MSG_FLAG_GPU(f)   (((f) >> 8) & 0xffu)

/* esignal.h:46,60-62 / fillet.h:76 values used by the thread bodies */
#define OTTGPU_SIGNAL_FRAME_REPEAT          0x11
#define OTTGPU_SIGNAL_DIRECT_ERROR_AVSYNC   0xe0
#define OTTGPU_SIGNAL_DIRECT_ERROR_MSGPOOL  0xe1
#define OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL  0xe2
#define OTTGPU_AV_SYNC_TRIGGER_LEVEL        30

#endif /* OTTGPU_FILLET_ABI_H_ */
