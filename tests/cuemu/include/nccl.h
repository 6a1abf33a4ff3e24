// nccl.h (cuemu) — TEST INFRASTRUCTURE.  Types only: solver.cu reaches NCCL through dlopen/dlsym, and the
// emulated library never creates a communicator (strips are exercised through the same-process transport).
#pragma once
#include <cuda_runtime.h>
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclInt64 = 4, ncclUint64 = 5, ncclFloat32 = 7, ncclFloat = 7 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
