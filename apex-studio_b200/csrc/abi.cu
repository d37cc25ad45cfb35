// Library identity + error strings of the C ABI (include/apex_b200.h).
#include "../../include/apex_b200.h"

extern "C" int b200_version(void) { return 100; }

extern "C" const char* b200_strerror(int code) {
  switch (code) {
    case B200_OK: return "ok";
    case B200_ERR_SHAPE: return "unsupported shape";
    case B200_ERR_ALIGN: return "pointer or stride is not 16-byte aligned";
    case B200_ERR_DRIVER: return "CUDA driver entry point unavailable (no GPU / driver?)";
    case B200_ERR_TMAP: return "cuTensorMapEncodeTiled rejected the tensor layout";
    case B200_ERR_LAUNCH: return "kernel launch failed";
    case B200_ERR_ARG: return "null pointer or invalid enum argument";
    default: return "unknown error";
  }
}
