// placeholder, replaced below
#include "tgpu_internal.h"
int cellrun_supported(const tgpu_ctx *) { return 0; }
int cellrun_move_deposit(tgpu_ctx *) { return TGPU_EINVAL; }
int cellrun_deposit(tgpu_ctx *) { return TGPU_EINVAL; }
