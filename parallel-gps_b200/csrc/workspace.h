// Handle = reusable device workspace + options.  Internal header (not part of the C ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>

enum { WS_LANE = 0, WS_WAGG, WS_WSTATE, WS_PART, WS_MISC, WS_COUNT };

struct pssgp_handle {
    int device;
    int num_sms;
    void* buf[WS_COUNT];
    size_t cap[WS_COUNT];
    unsigned int* ticket;
    int64_t chunk_opt;
    int64_t launches;
};
