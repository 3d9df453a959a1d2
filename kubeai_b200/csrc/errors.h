// Thread-local error string shared by all C-ABI translation units.
#pragma once

namespace b200 {
void set_error(const char* fmt, ...);
int require_device();            // B200_ERR_NO_DEVICE (+message) when no GPU is present
int cuda_fail(const char* what, int rc);
}  // namespace b200
