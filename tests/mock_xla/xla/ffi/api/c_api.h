/* Minimal stand-in for jaxlib's xla/ffi/api/c_api.h: only what enf_xla_ffi.cc touches, so that the shim type-checks in an
 * image without jaxlib (tests/test_cabi_cpu.py::test_xla_ffi_shim_type_checks).  NOT the real header: the real build uses
 * `python -c 'import jax.ffi; print(jax.ffi.include_dir())'`. */
#ifndef ENF_MOCK_XLA_FFI_C_API_H_
#define ENF_MOCK_XLA_FFI_C_API_H_
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_CallFrame XLA_FFI_CallFrame;
#endif
