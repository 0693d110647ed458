// Host planner: turns one channel configuration into device-resident polyphase tables, a tiling and a work-item list.
// Equivalent of the N sws_getContext() calls of video_scale_thread (transvideo.c:1562-1569) plus the thumbnail scaler
// (transvideo.c:2154-2158), done once per distinct configuration per GPU and shared by all channels that use it.
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
#include "../../include/ottgpu.h"
#include "device_types.h"
#include "polyphase.h"

namespace ottgpu {

struct RungLayout {                 // byte layout of one rung buffer
    int planes = 0;
    size_t offset[3] = {0, 0, 0};
    int pitch[3] = {0, 0, 0};
    size_t bytes = 0;
};

int fmt_bps(int fmt);                                  // 1 | 2, 0 if unknown
size_t packed_frame_bytes(int fmt, int w, int h);
bool rung_layout(const ottgpu_rung_cfg &r, RungLayout &out);

struct Plan {
    int gpu = 0;
    ottgpu_cfg cfg{};
    int n_slots = 0, thumb_slot = -1;
    RungLayout layout[kMaxRungSlots];
    PlanDev host_copy{};                               // what was uploaded (device pointers inside)
    PlanDev *dev = nullptr;
    std::vector<Item> items;                           // chan = 0; rung tiles inside the two-CTAs-per-SM budget, ordered by source row band
    std::vector<Item> big_items;                       // rung tiles of very steep / long filters: own (larger) launch
    std::vector<Item> thumb_items;                     // the thumbnail slot's items (launched only when due)
    // shared-memory region maxima (bytes) over the tiles of the launch classes: [0] items, [1] big_items, [2] thumbnail
    int src_max[3] = {0, 0, 0}, mid_max[3] = {0, 0, 0}, vc_max[3] = {0, 0, 0};
    std::vector<void *> allocations;                   // device memory owned by the plan
    ~Plan();
};

// returns nullptr and sets err on failure
std::shared_ptr<Plan> get_plan(const ottgpu_cfg &cfg, std::string &err);

// host-only: tables exactly as the plan builds them (ottgpu_filter_table, tests)
PolyphaseTable plan_table(int src_n, int dst_n, int filter, int target, bool vertical);

}  // namespace ottgpu
