/*
 * ottgpu.h -- C ABI of libottgpu: the B200 (sm_100a) replacement for the CPU video-preprocessing hot path of
 * cannonbeach/ott-packager's source/transvideo.c (yadif deinterlace -> N-rendition swscale ladder -> thumbnail).
 *
 * Plain C: pointers, ints and PODs only.  Every entry point names the reference interface it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the patch a maintainer applies to transvideo.c.
 *
 * Arithmetic contract: outputs are byte-identical to the reference's libswscale / libavfilter(yadif) calls at the
 * same flags.  OTTGPU_TARGET_A = what flags=SWS_BICUBIC (or SWS_LANCZOS) alone produce on
 * any x86 build of libswscale (the reference's flags, transvideo.c:1565); OTTGPU_TARGET_B = SWS_BITEXACT /
 * SWS_ACCURATE_RND / non-x86.
 *
 * There is NO CPU fallback: every call fails with OTTGPU_ERR_CUDA when no CUDA device / kernel image is usable.
 *
 * Threading (the reference runs one prepare/scale thread pair per channel, transvideo.c:3044-3099): any number of threads
 * may use the library at once; one object (channel, batch, converter) is driven by one thread at a time; pools are
 * internally locked and may be shared between threads (producer takes, consumer returns, like mempool.c).
 */
#ifndef OTTGPU_H_
#define OTTGPU_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define OTTGPU_ABI_VERSION 1
#define OTTGPU_MAX_RUNGS   8   /* == MAX_TRANS_OUTPUTS, include/fillet.h:57 */

/* pixel layouts (all 4:2:0) */
enum {
    OTTGPU_FMT_I420 = 0,       /* 8-bit planar Y,U,V        (AV_PIX_FMT_YUV420P, transvideo.c:1563-1564)            */
    OTTGPU_FMT_NV12 = 1,       /* 8-bit Y + interleaved UV  (NVENC sw_format, transvideo.c:403, repack 532-557)     */
    OTTGPU_FMT_P010 = 2,       /* 10-bit MSB-aligned in 16, Y + interleaved UV (little endian); the six low bits of a
                                  SOURCE sample must be zero, as decoders write them (libswscale shifts them out, the
                                  tensor-core H pass reads the raw bytes)                                            */
    OTTGPU_FMT_I420P10 = 3     /* 10-bit LSB-aligned in 16, planar (the only >8-bit layout vf_yadif accepts)        */
};
enum { OTTGPU_FILTER_BICUBIC = 0 /* SWS_BICUBIC, transvideo.c:1565 */, OTTGPU_FILTER_LANCZOS = 1 /* SWS_LANCZOS */ };
enum { OTTGPU_TARGET_A = 0, OTTGPU_TARGET_B = 1 };
enum {
    OTTGPU_DEINT_ALL = 0,      /* "yadif=0:-1:0"  transvideo.c:2045-2052 (every frame deinterlaced, 1-frame latency) */
    OTTGPU_DEINT_FLAGGED = 1,  /* "yadif=0:-1:1"  transvideo.c:2036-2044 (fps_num 60000|50000: only flagged frames) */
    OTTGPU_DEINT_BYPASS = 2    /* no deinterlacer stage at all, zero latency (plain sws_scale replacement)          */
};
enum { OTTGPU_MEM_HOST = 0, OTTGPU_MEM_DEVICE = 1,
       OTTGPU_MEM_HOST_WC = 2 /* ottgpu_pool_create only: pinned write-combined host memory for upload rings the CPU only
                                 writes (no cache snooping on the GPU's DMA reads); frames from it are OTTGPU_MEM_HOST */ };

/* error codes (functions return 0 on success, <0 on error; nothing in the library ever calls exit()) */
enum {
    OTTGPU_OK = 0,
    OTTGPU_ERR_ARG = -1,       /* bad argument / unsupported combination */
    OTTGPU_ERR_CUDA = -2,      /* CUDA runtime error, no device, no sm_100a image */
    OTTGPU_ERR_NOMEM = -3,     /* device / pinned allocation failed or pool exhausted */
    OTTGPU_ERR_STATE = -4      /* call sequence error */
};

typedef struct ottgpu_rung_cfg {
    int width, height;         /* core->cd->transvideo_info[i].width/height, transvideo.c:1557-1558 */
    int fmt;                   /* OTTGPU_FMT_*; bit depth must equal the source's */
    int pitch_align;           /* 0/1: rows packed like the reference's pool buffers (transvideo.c:1610-1624);
                                  N>1: every plane's row pitch rounded up to N bytes (NVENC-style surface) */
} ottgpu_rung_cfg;

typedef struct ottgpu_cfg {
    int gpu;                   /* CUDA device ordinal (core->cd->gpu, transvideo.c:327) */
    int src_width, src_height; /* msg->width/height, transvideo.c:1546-1547, 1955-1956 */
    int src_fmt;               /* OTTGPU_FMT_I420 | P010 | I420P10; packed rows (stride = width, transvideo.c:1549-1554) */
    int n_rungs;               /* core->cd->num_outputs */
    ottgpu_rung_cfg rung[OTTGPU_MAX_RUNGS];
    int filter;                /* OTTGPU_FILTER_* */
    int target;                /* OTTGPU_TARGET_* */
    int deint;                 /* OTTGPU_DEINT_* */
    int thumb_width, thumb_height; /* THUMBNAIL_WIDTH/HEIGHT 176x144, transvideo.c:157-158; 0 = no thumbnails */
    int thumb_period;          /* 150, transvideo.c:2160-2161 */
    int thumb_phase;           /* initial value of the thumbnail counter (0 in the reference); lets a multi-channel
                                  process stagger its channels' thumbnails */
    void *stream;              /* optional cudaStream_t to run on (NULL: the channel creates its own) */
} ottgpu_cfg;

typedef struct ottgpu_frame_in {
    const void *data;          /* msg->buffer: packed frame, Y then U then V (or Y then UV) */
    int mem;                   /* OTTGPU_MEM_HOST: uploaded by the call - complete on return for SYNC submits; with
                                  OTTGPU_SUBMIT_ASYNC the buffer must stay untouched until the step has retired
                                  (ottgpu_batch_wait, or the second-next submit of the batch has returned).
                                  OTTGPU_MEM_DEVICE: referenced in place, see released[] */
    int interlaced, tff;       /* msg->interlaced / msg->tff -> AVFrame flags, transvideo.c:2100-2101 */
    int64_t pts, dts;
    void *opaque;              /* rides with the frame through the 1-frame latency (AVFrame.opaque, transvideo.c:2103-2152) */
} ottgpu_frame_in;

typedef struct ottgpu_frame_out {
    /* ---- filled by the CALLER before the call ---- */
    void *dst[OTTGPU_MAX_RUNGS];     /* one buffer per rung, ottgpu_rung_bytes() large (memory_take(raw_video_pool), transvideo.c:1595) */
    int dst_mem[OTTGPU_MAX_RUNGS];   /* OTTGPU_MEM_DEVICE: written in place by the kernels (zero copy);  HOST: copied back */
    /* ---- filled by the LIBRARY ---- */
    int produced;              /* 1 if dst[] now holds a frame; 0 on the call that primes yadif (buffers untouched) */
    void *opaque;              /* opaque of the frame that was produced */
    int64_t pts, dts;          /* its timestamps (unscaled: the reference halves yadif's doubled pts back, transvideo.c:2325) */
    int interlaced, tff;       /* always 0 / 1 like transvideo.c:2327-2328 */
    void *thumbnail;           /* non-NULL every thumb_period-th produced frame: malloc()ed host I420 thumb_width x
                                  thumb_height in ONE allocation (Y, U, V; strides w, w/2, w/2) like av_image_alloc(align=1),
                                  transvideo.c:2165; the consumer free()s it (transvideo.c:254) */
    int n_released;            /* device-memory inputs the library no longer references (caller may recycle them) */
    const void *released[4];
} ottgpu_frame_out;

typedef struct ottgpu_channel ottgpu_channel;
typedef struct ottgpu_batch ottgpu_batch;

/* --- library -------------------------------------------------------------------------------------------------- */
int         ottgpu_abi_version(void);
const char *ottgpu_build_info(void);                    /* "libottgpu sm_100a ..." for the product build          */
int         ottgpu_device_count(void);                  /* <0 on error (no driver / no device)                   */
const char *ottgpu_strerror(int code);
const char *ottgpu_last_error(void);                    /* thread-local detail string of the last failure          */

/* --- geometry helpers (pure host arithmetic) -------------------------------------------------------------------- */
size_t ottgpu_frame_bytes(int fmt, int width, int height);         /* packed: 3*w*h/2 (x2 for 16-bit), transvideo.c:1559 */
size_t ottgpu_rung_bytes(const ottgpu_rung_cfg *r);                /* with pitch_align applied */
/* plane p (0..2; NV12/P010 have 2) of a rung buffer: byte offset from dst[i] and row pitch in bytes; returns planes */
int    ottgpu_rung_layout(const ottgpu_rung_cfg *r, size_t offset[3], int pitch[3]);

/* --- channel = one transvideo pipeline (replaces the avfilter graph + SwsContexts of one fillet process) ---------- */
/* create: sws_getContext x n_rungs (transvideo.c:1563) + avfilter_graph_* (2003-2054) + thumbnail scaler (2155)      */
int  ottgpu_channel_create(const ottgpu_cfg *cfg, ottgpu_channel **out);
/* submit: av_buffersrc_add_frame_flags + av_buffersink_get_frame (2132-2141) + sws_scale x n (1576) + thumbnail
 * sws_scale (2169) + the packed-buffer copies (1610-1624, 2217-2231) + NV12 repack (543-557).  Synchronous.            */
int  ottgpu_channel_submit(ottgpu_channel *ch, const ottgpu_frame_in *in, ottgpu_frame_out *out);
/* reset: source_discontinuity handling, transvideo.c:1959-1970: drops the yadif window (the frame submitted last is
 * never produced; fetch its opaque with ottgpu_channel_pending_opaque() first).  The thumbnail counter keeps running,
 * like the reference's thumbnail_count, which the discontinuity block does not touch.                                 */
int  ottgpu_channel_reset(ottgpu_channel *ch);
/* opaque of the frame held back by yadif's one-frame latency (submitted, not yet produced), or NULL.  Call it before
 * reset / destroy to release what rides with that frame (the reference leaks its opaque_struct there).               */
void *ottgpu_channel_pending_opaque(const ottgpu_channel *ch);
/* destroy: avfilter_graph_free / sws_freeContext, transvideo.c:1677, 1963-1966 */
void ottgpu_channel_destroy(ottgpu_channel *ch);
int  ottgpu_channel_gpu(const ottgpu_channel *ch);
/* human-readable plan of the channel (tables, tiling, shared memory, work items per frame) written to buf; returns
 * the number of bytes that would have been written (snprintf semantics) */
int  ottgpu_channel_describe(const ottgpu_channel *ch, char *buf, size_t size);

/* --- batch = many channels of one GPU advanced by ONE fused launch per stage (in-process multi-channel daemon) ----- */
enum { OTTGPU_SUBMIT_SYNC = 0, OTTGPU_SUBMIT_ASYNC = 1 };
int  ottgpu_batch_create(ottgpu_channel **channels, int n, ottgpu_batch **out);    /* channels must share gpu + stream */
/* in[n] / out[n]; in[i].data == NULL skips channel i this step.  ASYNC: returns after enqueueing; at most two steps are
 * in flight (a third submit first waits for the oldest).  Host-memory results, out[] fields and the CONTENT of
 * out[i].thumbnail are valid once the step has completed: after ottgpu_batch_wait(), or when the second-next submit of
 * the same batch has returned.  A thumbnail buffer must not be free()d before that; host INPUT frames must not be
 * overwritten before that either (their upload is asynchronous).  A failed submit leaves the batch and its channels
 * exactly as they were before the call (windows, slots, thumbnail counters), out[].thumbnail NULL, nothing leaked.    */
int  ottgpu_batch_submit(ottgpu_batch *b, const ottgpu_frame_in *in, ottgpu_frame_out *out, int flags);
int  ottgpu_batch_wait(ottgpu_batch *b);
void ottgpu_batch_destroy(ottgpu_batch *b);
/* counters for bench.py: kernels launched / work items (CTAs) in the last submit */
int  ottgpu_batch_last_launches(const ottgpu_batch *b);
int  ottgpu_batch_last_items(const ottgpu_batch *b);
/* host-to-device frame copies the last submit issued: consecutive buffers of one pinned ottgpu_pool going to channels of
 * one batch travel as ONE copy per step */
int  ottgpu_batch_last_uploads(const ottgpu_batch *b);
/* device-to-host copies the last submit issued for host-destined rungs and thumbnails: the surfaces of one rung of all
 * channels live in one device allocation, so channels whose destinations are consecutive buffers of one pinned
 * ottgpu_pool are read back with ONE copy per rung (the x264/x265 hand-off of transvideo.c:1610-1624 for N channels) */
int  ottgpu_batch_last_readbacks(const ottgpu_batch *b);
/* device-side timing of the ladder launches (CUDA events recorded on the batch's stream around each launch).
 * enable != 0 starts/reset accumulation; read (after ottgpu_batch_wait) returns the summed milliseconds and count. */
int  ottgpu_batch_timing(ottgpu_batch *b, int enable);
int  ottgpu_batch_timing_read(ottgpu_batch *b, double *ms_sum, int *launches);
int  ottgpu_batch_timing_read_yadif(ottgpu_batch *b, double *ms_sum, int *launches);   /* same for the yadif launches */

/* --- pools: the mempool.h discipline (memory_create / memory_take / memory_return / memory_unused,
 *     include/mempool.h:30-35, source/mempool.c:82-265) for device surfaces and pinned host frames ------------------ */
typedef struct ottgpu_pool ottgpu_pool;
int    ottgpu_pool_create(int gpu, int mem, int buffer_count, size_t buffer_size, ottgpu_pool **out);
void  *ottgpu_pool_take(ottgpu_pool *p);                /* NULL when exhausted (memory_take semantics)             */
int    ottgpu_pool_return(ottgpu_pool *p, void *buffer);
int    ottgpu_pool_unused(ottgpu_pool *p);
void   ottgpu_pool_destroy(ottgpu_pool *p);
/* plain synchronous copy between host and device buffers (the encoder-side hand-off of a device rung to a CPU encoder,
 * i.e. what replaces the memcpy of transvideo.c:1330-1338 when the rung lives in HBM) */
int    ottgpu_memcpy(int gpu, void *dst, int dst_mem, const void *src, int src_mem, size_t bytes);

/* --- decode-side converter ------------------------------------------------------------------------------------------
 * Replaces the decode thread's `decode_converter` (transvideo.c:2797-2808: sws_getContext(w,h,source_format -> w,h,
 * AV_PIX_FMT_YUV420P, SWS_BICUBIC) + sws_scale) together with the three row-copy loops that pack the planes into the
 * contiguous w*h*3/2 frame handed to the prepare thread (transvideo.c:2823-2840).  The result is byte-identical to
 * libswscale's for the same flags (`target` as in ottgpu_cfg).  With dst_mem = OTTGPU_MEM_DEVICE the packed frame stays
 * on the GPU and is fed to ottgpu_channel_submit / ottgpu_batch_submit as an OTTGPU_MEM_DEVICE input. */
enum {
    OTTGPU_PIX_YUV420P   = 0,   /* AV_PIX_FMT_YUV420P: no arithmetic, rows re-packed                                */
    OTTGPU_PIX_YUV422P   = 1,   /* AV_PIX_FMT_YUV422P                                                               */
    OTTGPU_PIX_YUV420P10 = 2,   /* AV_PIX_FMT_YUV420P10LE (samples 0..1023 in 16-bit words)                         */
    OTTGPU_PIX_YUV422P10 = 3    /* AV_PIX_FMT_YUV422P10LE                                                           */
};

typedef struct ottgpu_picture {
    const void *plane[3];       /* Y, U, V as the decoder hands them over (AVFrame.data, transvideo.c:2788-2791)    */
    int         stride[3];      /* bytes per row (AVFrame.linesize, transvideo.c:2792-2795); >= the row's bytes     */
    int         mem;            /* OTTGPU_MEM_HOST or OTTGPU_MEM_DEVICE                                             */
} ottgpu_picture;

typedef struct ottgpu_converter ottgpu_converter;

/* stream: the caller's cudaStream_t (work is ordered on it) or NULL for a private stream */
int  ottgpu_converter_create(int gpu, int pix_fmt, int width, int height, int target, void *stream, ottgpu_converter **out);
/* dst: packed I420 frame, ottgpu_frame_bytes(OTTGPU_FMT_I420, width, height) bytes, in dst_mem.
 * flags: 0 = the frame is complete on return; OTTGPU_SUBMIT_ASYNC = enqueued on the stream (host pictures must stay
 * valid until ottgpu_converter_wait / a stream synchronisation) */
int  ottgpu_converter_convert(ottgpu_converter *cv, const ottgpu_picture *in, void *dst, int dst_mem, int flags);
/* n pictures of n converters (one per channel, same gpu) in ONE launch, ordered on cv[0]'s stream: a 1080p conversion is
 * launch-bound, a multi-channel process converts all of a step's pictures together.  in[i] / dst[i] belong to cv[i]. */
int  ottgpu_converter_convert_batch(ottgpu_converter *const *cv, const ottgpu_picture *in, void *const *dst, int n, int dst_mem, int flags);
int  ottgpu_converter_wait(ottgpu_converter *cv);
void ottgpu_converter_destroy(ottgpu_converter *cv);

/* --- stateless single-shot helpers (unit tests, golden comparisons; same kernels) --------------------------------- */
/* sws_getContext + sws_scale on one packed host frame: src (fmt,w,h) -> dst rung; dst is ottgpu_rung_bytes() large   */
int ottgpu_scale_frame(int gpu, const void *src, int src_fmt, int src_w, int src_h,
                       void *dst, const ottgpu_rung_cfg *rung, int filter, int target);
/* one yadif output frame from packed host frames prev/cur/next (I420 or I420P10) */
int ottgpu_yadif_frame(int gpu, const void *prev, const void *cur, const void *next, void *dst,
                       int fmt, int width, int height, int interlaced, int tff);
/* the polyphase table sws_getContext would build: returns taps; pos[dst_n], coef[dst_n*taps] (caller sizes coef for
 * dst_n*256).  vertical selects V (12-bit) vs H (14-bit) conventions */
int ottgpu_filter_table(int src_n, int dst_n, int filter, int target, int vertical, int32_t *pos, int16_t *coef);

/* ------------------------------------------------------------------------------------------------------------------
 * Thumbnail JPEG on the GPU.  Replaces save_frame_as_jpeg() of video_thumbnail_thread
 * (source/transvideo.c:164-198: avcodec_find_encoder(AV_CODEC_ID_MJPEG), 176x144 YUVJ420P, one frame per call).
 * Baseline JPEG, 4:2:0, planes taken as Y/Cb/Cr as they are, ONE quantisation table for all components built the way
 * libavcodec's MJPEG encoder builds it: q[i] = clip((ff_mpeg1_default_intra_matrix[i] * qscale) >> 3, 1, 255), q[0] = 8
 * (it picks qscale 2-3 for natural content at the reference's 250 kb/s), and its quantiser arithmetic (3/8 intra bias, DC
 * scale 8): every DC level and > 99.5 % of the AC levels equal the real encoder's; like libavcodec the picture is coded
 * with its own optimal Huffman tables (same segment sequence COM DQT DHT SOF0 SOS, file size within 2 % of its file).
 * The forward DCT (IJG integer DCT) and the table construction are this library's own: the files decode to the same
 * picture as the reference's, they are not byte-identical (no bit-exact oracle for its SIMD DCT / rate control).
 */
typedef struct ottgpu_jpeg ottgpu_jpeg;
int  ottgpu_jpeg_create(int gpu, int width, int height, int qscale, ottgpu_jpeg **out);
/* pic: three planes + strides in host or device memory (the scale_struct of the thumbnail hand-off, transvideo.c:2176-2183).
 * dst: host buffer of dst_capacity bytes; *dst_bytes = size of the file.  Synchronous.  OTTGPU_ERR_NOMEM when dst is too
 * small (ottgpu_jpeg_max_bytes() is always enough). */
int  ottgpu_jpeg_encode(ottgpu_jpeg *j, const ottgpu_picture *pic, void *dst, size_t dst_capacity, size_t *dst_bytes);
size_t ottgpu_jpeg_max_bytes(const ottgpu_jpeg *j);
void ottgpu_jpeg_destroy(ottgpu_jpeg *j);

#if defined(__cplusplus)
}
#endif
#endif /* OTTGPU_H_ */
