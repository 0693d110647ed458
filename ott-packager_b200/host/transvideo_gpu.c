/* See transvideo_gpu.h.  Control flow follows source/transvideo.c:1896-2429 (prepare) and 1504-1683 (scale); the
 * libavfilter / libswscale calls and the per-plane memcpy loops of those functions are replaced by one
 * ottgpu_channel_submit().  Per-frame fprintf(stderr) of the reference is intentionally not reproduced. */
#include "transvideo_gpu.h"
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* opaque_struct of transvideo.c:129-136: sideband that rides with a frame through yadif's one-frame latency */
typedef struct tv_sideband {
    uint8_t *caption_buffer;
    int      caption_size;
    int64_t  caption_timestamp;
    int64_t  splice_duration;
    int      splice_point;
    int64_t  splice_duration_remaining;
    int64_t  pts, dts;
    int      fps_num, fps_den, aspect_num, aspect_den;
} tv_sideband;

static void tv_fatal(ottgpu_transvideo *tv, int code, const char *msg)
{
    if (tv->direct_error) tv->direct_error(tv->user, code, msg);
}

/* submitted: the channel has been handed the frame.  Host frames were copied during that call; device frames stay
 * referenced by the deinterlacer window and come back through ottgpu_frame_out.released[] (tv_settle_inputs). */
static void tv_return_input(ottgpu_transvideo *tv, dataqueue_message_struct *msg, int submitted)
{
    /* transvideo.c:2410-2413 */
    if (msg->flags & OTTGPU_MSG_FLAG_DEVICE) {
        if (!submitted && msg->buffer) ottgpu_pool_return(tv->input_pool, msg->buffer);
    } else {
        memory_return(tv->raw_video_pool, msg->buffer);
    }
    msg->buffer = NULL;
    memory_return(tv->msg_pool, msg);
}

/* the frame yadif still holds back never comes out after a reset / re-key / close: free what rides with it */
static void tv_drop_pending(ottgpu_transvideo *tv)
{
    tv_sideband *sb;
    if (!tv->channel) return;
    sb = (tv_sideband *)ottgpu_channel_pending_opaque(tv->channel);
    if (sb) {
        free(sb->caption_buffer);
        free(sb);
    }
}

static void tv_release_held(ottgpu_transvideo *tv)
{
    int i;
    for (i = 0; i < tv->n_held; ++i) ottgpu_pool_return(tv->input_pool, tv->held_inputs[i]);
    tv->n_held = 0;
}

/* after a submit: remember the device frame just handed over, give back the ones the channel let go of */
static void tv_settle_inputs(ottgpu_transvideo *tv, void *submitted_device_frame, const ottgpu_frame_out *out)
{
    int r, i;
    if (submitted_device_frame && tv->n_held < (int)(sizeof(tv->held_inputs) / sizeof(tv->held_inputs[0])))
        tv->held_inputs[tv->n_held++] = submitted_device_frame;
    for (r = 0; r < out->n_released; ++r)
        for (i = 0; i < tv->n_held; ++i)
            if (tv->held_inputs[i] == out->released[r]) {
                ottgpu_pool_return(tv->input_pool, tv->held_inputs[i]);
                tv->held_inputs[i] = tv->held_inputs[--tv->n_held];
                break;
            }
}

static int tv_open_channel(ottgpu_transvideo *tv, const dataqueue_message_struct *msg)
{
    ottgpu_cfg cfg;
    /* "yadif=0:-1:1" only for 50/60 fps sources, else blanket deinterlace (transvideo.c:2036-2052) */
    const int deint = (msg->fps_num == 60000 || msg->fps_num == 50000) ? OTTGPU_DEINT_FLAGGED : OTTGPU_DEINT_ALL;
    if (tv->channel && tv->chan_width == msg->width && tv->chan_height == msg->height && tv->chan_deint == deint) return 0;
    tv_drop_pending(tv);
    if (tv->channel) ottgpu_channel_destroy(tv->channel);   /* the reference never re-keys its scalers (latent bug) */
    tv->channel = NULL;
    tv_release_held(tv);
    memset(&cfg, 0, sizeof(cfg));
    cfg.gpu = tv->gpu;
    cfg.src_width = msg->width;
    cfg.src_height = msg->height;
    cfg.src_fmt = OTTGPU_FMT_I420;                           /* AV_PIX_FMT_YUV420P, transvideo.c:1910 */
    cfg.n_rungs = tv->num_outputs;
    for (int i = 0; i < tv->num_outputs; ++i) {
        cfg.rung[i].width = tv->out_width[i];
        cfg.rung[i].height = tv->out_height[i];
        cfg.rung[i].fmt = tv->out_fmt;
        cfg.rung[i].pitch_align = 0;                         /* packed, transvideo.c:1610-1624 */
    }
    cfg.filter = OTTGPU_FILTER_BICUBIC;                      /* SWS_BICUBIC, transvideo.c:1565 */
    cfg.target = tv->target;
    cfg.deint = deint;
    cfg.thumb_width = 176;                                   /* THUMBNAIL_WIDTH / HEIGHT, transvideo.c:157-158 */
    cfg.thumb_height = 144;
    cfg.thumb_period = 150;                                  /* transvideo.c:2160-2161 */
    if (ottgpu_channel_create(&cfg, &tv->channel) != OTTGPU_OK) return -1;
    tv->chan_width = msg->width;
    tv->chan_height = msg->height;
    tv->chan_deint = deint;
    return 0;
}

static void *tv_take_rung(ottgpu_transvideo *tv, int i, size_t bytes)
{
    return tv->device_outputs ? ottgpu_pool_take(tv->surface_pool[i]) : memory_take(tv->raw_video_pool, (int)bytes);
}

static void tv_return_rung(ottgpu_transvideo *tv, int i, void *p)
{
    if (!p) return;
    if (tv->device_outputs) ottgpu_pool_return(tv->surface_pool[i], p);
    else memory_return(tv->raw_video_pool, p);
}

/* transvideo.c:1626-1664: one message per rendition */
static int tv_emit(ottgpu_transvideo *tv, int i, void *buffer, size_t bytes, const tv_sideband *sb, int last)
{
    dataqueue_message_struct *m = (dataqueue_message_struct *)memory_take(tv->msg_pool, sizeof(dataqueue_message_struct));
    if (!m) {
        tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_MSGPOOL, "Out of Message Buffers (SCALE) - Restarting Service");
        return -1;
    }
    m->buffer = buffer;
    m->buffer_size = bytes;
    m->flags = tv->device_outputs ? (OTTGPU_MSG_FLAG_DEVICE | ((uint32_t)tv->gpu << 8)) : 0;
    m->pts = sb->pts;
    m->dts = sb->dts;
    m->interlaced = 0;
    m->tff = 1;
    m->fps_num = sb->fps_num;
    m->fps_den = sb->fps_den;
    m->aspect_num = sb->aspect_num;
    m->aspect_den = sb->aspect_den;
    m->width = tv->out_width[i];
    m->height = tv->out_height[i];
    m->stream_index = (uint8_t)i;
    m->splice_point = sb->splice_point;
    m->splice_duration = sb->splice_duration;
    m->splice_duration_remaining = sb->splice_duration_remaining;
    m->caption_size = sb->caption_size;
    m->caption_timestamp = sb->caption_timestamp;
    m->caption_buffer = NULL;
    if (sb->caption_size > 0 && sb->caption_buffer) {        /* last rendition takes the buffer, the others copies */
        if (last) {
            m->caption_buffer = sb->caption_buffer;
        } else {
            m->caption_buffer = (uint8_t *)malloc((size_t)sb->caption_size);
            if (m->caption_buffer) memcpy(m->caption_buffer, sb->caption_buffer, (size_t)sb->caption_size);
        }
    }
    dataqueue_put_front(tv->encode_queue[i], m);
    return 0;
}

int ottgpu_transvideo_process(ottgpu_transvideo *tv, dataqueue_message_struct *msg)
{
    ottgpu_frame_in in;
    ottgpu_frame_out out;
    size_t bytes[OTTGPU_MAX_RUNGS];
    tv_sideband *sb;
    int i, rc;

    tv->frames_in++;
    if (tv_open_channel(tv, msg) != 0) {
        tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL, ottgpu_last_error());
        tv_return_input(tv, msg, 0);
        return -1;
    }
    if (msg->source_discontinuity) {                         /* transvideo.c:1959-1970 */
        tv_drop_pending(tv);
        ottgpu_channel_reset(tv->channel);
        tv_release_held(tv);
        msg->source_discontinuity = 0;
    }
    /* sideband for this frame (transvideo.c:2103-2128); it comes back with the frame it belongs to */
    sb = (tv_sideband *)calloc(1, sizeof(*sb));
    if (!sb) { tv_return_input(tv, msg, 0); return -1; }
    if (msg->caption_buffer) {
        sb->caption_buffer = msg->caption_buffer;
        sb->caption_size = msg->caption_size;
        sb->caption_timestamp = msg->caption_timestamp;
    }
    sb->splice_point = msg->splice_point;
    sb->splice_duration = msg->splice_duration;
    sb->splice_duration_remaining = msg->splice_duration_remaining;
    sb->pts = msg->pts;
    sb->dts = msg->dts;
    sb->fps_num = msg->fps_num; sb->fps_den = msg->fps_den;
    sb->aspect_num = msg->aspect_num; sb->aspect_den = msg->aspect_den;

    memset(&in, 0, sizeof(in));
    memset(&out, 0, sizeof(out));
    in.data = msg->buffer;
    in.mem = (msg->flags & OTTGPU_MSG_FLAG_DEVICE) ? OTTGPU_MEM_DEVICE : OTTGPU_MEM_HOST;
    in.interlaced = msg->interlaced;                         /* transvideo.c:2100-2101 */
    in.tff = msg->tff;
    in.pts = msg->pts;
    in.dts = msg->dts;
    in.opaque = sb;
    for (i = 0; i < tv->num_outputs; ++i) {
        ottgpu_rung_cfg rc_;
        rc_.width = tv->out_width[i]; rc_.height = tv->out_height[i]; rc_.fmt = tv->out_fmt; rc_.pitch_align = 0;
        bytes[i] = ottgpu_rung_bytes(&rc_);
        out.dst[i] = tv_take_rung(tv, i, bytes[i]);
        out.dst_mem[i] = tv->device_outputs ? OTTGPU_MEM_DEVICE : OTTGPU_MEM_HOST;
        if (!out.dst[i]) {
            int k;
            tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL, "Out of Uncompressed Video Buffers (SCALE) - Restarting Service");
            for (k = 0; k < i; ++k) tv_return_rung(tv, k, out.dst[k]);
            free(sb);
            tv_return_input(tv, msg, 0);
            return -1;
        }
    }
    rc = ottgpu_channel_submit(tv->channel, &in, &out);
    if (rc == OTTGPU_OK) tv_settle_inputs(tv, in.mem == OTTGPU_MEM_DEVICE ? msg->buffer : NULL, &out);
    if (rc != OTTGPU_OK || !out.produced) {
        for (i = 0; i < tv->num_outputs; ++i) tv_return_rung(tv, i, out.dst[i]);
        if (rc != OTTGPU_OK) {
            /* a failed submit leaves the channel as it was (the frame was not taken into the window), so this frame's
             * sideband is ours to free; the reference exit(0)s here, the hook decides */
            tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL, ottgpu_last_error());
            free(sb->caption_buffer);
            free(sb);
        }
        tv_return_input(tv, msg, rc == OTTGPU_OK);           /* the host copy of the frame is no longer needed */
        return rc == OTTGPU_OK ? 0 : -1;
    }

    {
        tv_sideband *osb = (tv_sideband *)out.opaque;        /* sideband of the frame that came out (one frame older) */
        double fps = osb->fps_den > 0 ? (double)osb->fps_num / (double)osb->fps_den : 30.0;
        int64_t sync_diff = 0;
        int drop = 0, repeat = 0;

        /* thumbnail hand-off (transvideo.c:2176-2186): scale_struct + single allocation, consumer frees both */
        if (out.thumbnail) {
            dataqueue_message_struct *tm = (dataqueue_message_struct *)memory_take(tv->msg_pool, sizeof(dataqueue_message_struct));
            ottgpu_scale_struct *ts = (ottgpu_scale_struct *)calloc(1, sizeof(*ts));
            if (!tm || !ts) {
                tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_MSGPOOL, "Out of Message Buffers (Thumbnail) - Restarting Service");
                free(out.thumbnail);
                free(ts);
            } else {
                uint8_t *t = (uint8_t *)out.thumbnail;
                ts->output_data[0] = t;
                ts->output_data[1] = t + 176 * 144;
                ts->output_data[2] = t + 176 * 144 + 88 * 72;
                ts->output_stride[0] = 176; ts->output_stride[1] = 88; ts->output_stride[2] = 88;
                tm->buffer = ts;
                dataqueue_put_front(tv->thumbnail_queue, tm);
                tv->thumbnails_out++;
            }
        }

        /* A/V sync accounting (transvideo.c:2233-2269): frames since start vs frames the timestamps imply */
        tv->deinterlaced_frame_count++;
        if (tv->first_timestamp) {
            int64_t sync_frame_count = (int64_t)((((double)osb->dts - (double)*tv->first_timestamp) / 90000.0) * fps);
            sync_diff = tv->deinterlaced_frame_count - sync_frame_count;
            if (sync_diff > (2 * fps) || sync_diff < (-fps * 2)) {
                tv->av_sync_compromised++;
                if (tv->av_sync_compromised == 1) tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_AVSYNC, "Possible A/V Sync Drift, Check Input");
                if (tv->av_sync_compromised >= OTTGPU_AV_SYNC_TRIGGER_LEVEL) {
                    tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_AVSYNC, "A/V Sync Is Compromised (VIDEO) - Restarting Service");
                    for (i = 0; i < tv->num_outputs; ++i) tv_return_rung(tv, i, out.dst[i]);
                    free(osb->caption_buffer);
                    free(osb);
                    tv_return_input(tv, msg, 1);
                    return -1;
                }
            } else {
                tv->av_sync_compromised = 0;
            }
            repeat = sync_diff < -1;                         /* transvideo.c:2271 */
            drop = sync_diff > 1 && (osb->splice_point == 0 || osb->splice_duration_remaining == 0);   /* 2366-2382 */
        }

        if (repeat) {
            /* repeat the frame to maintain A/V sync (transvideo.c:2271-2316): same pictures, pts/dts 0, no captions.
             * Device surfaces are duplicated on the device (no host round trip). */
            tv_sideband rsb = *osb;
            rsb.pts = rsb.dts = 0;
            rsb.caption_buffer = NULL;
            rsb.caption_size = 0;
            rsb.caption_timestamp = 0;
            if (tv->signal) tv->signal(tv->user, OTTGPU_SIGNAL_FRAME_REPEAT, "Repeating Video Frame To Maintain A/V Sync");
            for (i = 0; i < tv->num_outputs; ++i) {
                void *copy = tv_take_rung(tv, i, bytes[i]);
                if (!copy) {
                    tv_fatal(tv, OTTGPU_SIGNAL_DIRECT_ERROR_RAWPOOL, "Out of Uncompressed Video Buffers (SYNC) - Restarting Service");
                    break;
                }
                if (tv->device_outputs) ottgpu_memcpy(tv->gpu, copy, OTTGPU_MEM_DEVICE, out.dst[i], OTTGPU_MEM_DEVICE, bytes[i]);
                else memcpy(copy, out.dst[i], bytes[i]);
                if (tv_emit(tv, i, copy, bytes[i], &rsb, 0) != 0) { tv_return_rung(tv, i, copy); break; }
            }
            tv->deinterlaced_frame_count++;
        }

        if (drop) {
            /* dropped to catch up: captions of the dropped frame are discarded like the reference does (2368-2373) */
            for (i = 0; i < tv->num_outputs; ++i) tv_return_rung(tv, i, out.dst[i]);
            free(osb->caption_buffer);
            tv->deinterlaced_frame_count--;
        } else {
            int emitted_caption = 0;
            for (i = 0; i < tv->num_outputs; ++i) {
                const int last = i == tv->num_outputs - 1;
                if (tv_emit(tv, i, out.dst[i], bytes[i], osb, last) != 0) {
                    int k;
                    for (k = i; k < tv->num_outputs; ++k) tv_return_rung(tv, k, out.dst[k]);
                    break;
                }
                if (last) emitted_caption = 1;
            }
            if (!emitted_caption || osb->caption_size <= 0) free(osb->caption_buffer);
            tv->frames_out++;
        }
        free(osb);
    }
    tv_return_input(tv, msg, 1);
    return 0;
}

void *ottgpu_video_prepare_scale_thread(void *context)
{
    ottgpu_transvideo *tv = (ottgpu_transvideo *)context;
    while (*tv->running) {
        dataqueue_message_struct *msg = dataqueue_take_back(tv->prepare_queue);
        while (!msg && *tv->running) {                       /* transvideo.c:1935-1939 */
            usleep(1000);
            msg = dataqueue_take_back(tv->prepare_queue);
        }
        if (!*tv->running) {                                 /* drain, transvideo.c:1940-1952 */
            while (msg) {
                tv_return_input(tv, msg, 0);
                msg = dataqueue_take_back(tv->prepare_queue);
            }
            break;
        }
        if (msg && ottgpu_transvideo_process(tv, msg) != 0 && !*tv->running) break;
    }
    ottgpu_transvideo_close(tv);
    return NULL;
}

void ottgpu_transvideo_close(ottgpu_transvideo *tv)
{
    tv_drop_pending(tv);
    if (tv->channel) ottgpu_channel_destroy(tv->channel);    /* avfilter_graph_free / sws_freeContext, 2416-2427 */
    tv->channel = NULL;
    tv_release_held(tv);
}
