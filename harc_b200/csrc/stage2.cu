// placeholder, replaced below
#include "ctx.h"
int s2_set_stream_from_stage1(harcgpu_ctx *c) { harcgpu_set_error("stage II not built"); return -1; }
int s2_set_stream_host(harcgpu_ctx *c, const char *, const char *, const u8 *, const u32 *, const char *, u32) { harcgpu_set_error("stage II not built"); return -1; }
int s2_load_pool(harcgpu_ctx *c, const char *, const u32 *, u32, const char *, u32) { harcgpu_set_error("stage II not built"); return -1; }
int s2_encode(harcgpu_ctx *c) { harcgpu_set_error("stage II not built"); return -1; }
