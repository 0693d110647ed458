/* See encode_adapter.h. */
#include "encode_adapter.h"
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int is_device(const dataqueue_message_struct *msg)
{
    return (msg->flags & 0xffff00ffu) == OTTGPU_MSG_FLAG_DEVICE;
}

static int layout(const ottgpu_encode_adapter *ad, size_t offset[3], int pitch[3], size_t *bytes)
{
    ottgpu_rung_cfg r;
    r.width = ad->width; r.height = ad->height; r.fmt = ad->fmt; r.pitch_align = ad->pitch_align;
    *bytes = ottgpu_rung_bytes(&r);
    return ottgpu_rung_layout(&r, offset, pitch);
}

int ottgpu_encode_map_host(ottgpu_encode_adapter *ad, const dataqueue_message_struct *msg, ottgpu_encode_planes *out)
{
    size_t offset[3] = {0, 0, 0}, bytes = 0;
    int pitch[3] = {0, 0, 0}, i;
    const uint8_t *base;
    const int planes = layout(ad, offset, pitch, &bytes);
    if (planes <= 0 || !msg || !msg->buffer || !out) return OTTGPU_ERR_ARG;
    if (is_device(msg)) {
        if (!ad->staging || ad->staging_bytes < bytes) {                 /* pinned: the D2H runs at link speed */
            ottgpu_encode_adapter_close(ad);
            if (ottgpu_pool_create(ad->gpu, OTTGPU_MEM_HOST, 1, bytes, &ad->staging_pool) != OTTGPU_OK) return OTTGPU_ERR_NOMEM;
            ad->staging = ottgpu_pool_take(ad->staging_pool);
            ad->staging_bytes = bytes;
        }
        if (ottgpu_memcpy(ad->gpu, ad->staging, OTTGPU_MEM_HOST, msg->buffer, OTTGPU_MEM_DEVICE, bytes) != OTTGPU_OK) return OTTGPU_ERR_CUDA;
        base = (const uint8_t *)ad->staging;
        ad->d2h_frames++;
    } else {
        base = (const uint8_t *)msg->buffer;                             /* transvideo.c:962-967: pointers into the pool buffer */
    }
    out->planes = planes;
    for (i = 0; i < 3; ++i) {
        out->plane[i] = i < planes ? base + offset[i] : NULL;
        out->stride[i] = i < planes ? pitch[i] : 0;
    }
    ad->frames++;
    return OTTGPU_OK;
}

int ottgpu_encode_nvenc_surface(const ottgpu_encode_adapter *ad, const dataqueue_message_struct *msg, ottgpu_nvenc_surface *out)
{
    size_t offset[3] = {0, 0, 0}, bytes = 0;
    int pitch[3] = {0, 0, 0};
    if (!msg || !out || !is_device(msg)) return OTTGPU_ERR_ARG;          /* host buffers take the reference's own path */
    if (ad->fmt != OTTGPU_FMT_NV12 && ad->fmt != OTTGPU_FMT_P010) return OTTGPU_ERR_ARG;
    if (layout(ad, offset, pitch, &bytes) != 2 || pitch[0] != pitch[1]) return OTTGPU_ERR_ARG;
    out->device_ptr = msg->buffer;
    out->chroma_offset = offset[1];
    out->pitch = pitch[0];
    out->width = ad->width;
    out->height = ad->height;
    out->fmt = ad->fmt;
    out->gpu = (int)OTTGPU_MSG_FLAG_GPU(msg->flags);
    return OTTGPU_OK;
}

void ottgpu_encode_release(ottgpu_encode_adapter *ad, dataqueue_message_struct *msg)
{
    if (!msg) return;
    if (msg->buffer) {
        if (is_device(msg)) ottgpu_pool_return(ad->surface_pool, msg->buffer);
        else memory_return(ad->raw_video_pool, msg->buffer);
        msg->buffer = NULL;
    }
    memory_return(ad->msg_pool, msg);
}

int ottgpu_nvenc_probe(char *buf, size_t size)
{
    typedef int (*max_version_fn)(unsigned int *);
    void *h = dlopen("libnvidia-encode.so.1", RTLD_LAZY | RTLD_LOCAL);
    if (!h) {
        if (buf && size) snprintf(buf, size, "libnvidia-encode.so.1: not present (%s)", dlerror());
        return 0;
    }
    {
        unsigned int v = 0;
        max_version_fn fn = (max_version_fn)dlsym(h, "NvEncodeAPIGetMaxSupportedVersion");
        const int rc = fn ? fn(&v) : -1;
        if (buf && size)
            snprintf(buf, size, "libnvidia-encode.so.1: present, NvEncodeAPIGetMaxSupportedVersion rc=%d version=%u.%u "
                                "(an encode SESSION still needs NVENC hardware on the device)", rc, v >> 4, v & 15);
        dlclose(h);
        return fn && rc == 0;
    }
}

void ottgpu_encode_adapter_close(ottgpu_encode_adapter *ad)
{
    if (ad->staging_pool) {
        if (ad->staging) ottgpu_pool_return(ad->staging_pool, ad->staging);
        ottgpu_pool_destroy(ad->staging_pool);
    }
    ad->staging_pool = NULL;
    ad->staging = NULL;
    ad->staging_bytes = 0;
}
