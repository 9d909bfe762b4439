/* Minimal DLPack v0.x ABI structs (the public DLPack standard layout), declared here so
 * the C-ABI in cnf.h needs no third-party header.  Tensors cross the boundary as
 * `DLManagedTensor*` obtained zero-copy from the producer (torch.utils.dlpack.to_dlpack);
 * libcnf only BORROWS them for the duration of a call and never invokes the deleter.
 */
#ifndef CNF_DLPACK_H_
#define CNF_DLPACK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DLPACK_VERSION   /* do not clash with a real dlpack.h if one is included first */

typedef enum {
  kDLCPU = 1,
  kDLCUDA = 2,
  kDLCUDAHost = 3,
  kDLCUDAManaged = 13
} DLDeviceType;

typedef struct {
  DLDeviceType device_type;
  int32_t device_id;
} DLDevice;

typedef enum {
  kDLInt = 0U,
  kDLUInt = 1U,
  kDLFloat = 2U,
  kDLOpaqueHandle = 3U,
  kDLBfloat = 4U
} DLDataTypeCode;

typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} DLDataType;

typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL means compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;

#endif /* DLPACK_VERSION */

#ifdef __cplusplus
}
#endif
#endif /* CNF_DLPACK_H_ */
