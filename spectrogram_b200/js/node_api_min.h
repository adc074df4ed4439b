/*
 * node_api_min.h -- the subset of Node-API (N-API, ABI-stable since Node 8) that napi_shim.c uses.
 *
 * Node's own <node_api.h> is not present in this image (no Node toolchain), so the prototypes are
 * declared here by hand.  Node-API is a C ABI resolved from the host process when the addon is
 * loaded: an addon only needs these declarations at compile time and leaves the napi_* symbols
 * undefined in the shared object.  With a real Node installation, replace this header by
 * #include <node_api.h>; nothing else changes.
 */
#ifndef NODE_API_MIN_H_
#define NODE_API_MIN_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;

typedef enum {
  napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
  napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
  napi_generic_failure, napi_pending_exception, napi_cancelled, napi_escape_called_twice
} napi_status;

typedef enum {
  napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object,
  napi_function, napi_external, napi_bigint
} napi_valuetype;

typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
  napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array,
  napi_biguint64_array
} napi_typedarray_type;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);

napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb,
                                 void* data, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_has_named_property(napi_env env, napi_value object, const char* utf8name, bool* result);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv,
                             napi_value* this_arg, void** data);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t* result);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char* buf, size_t bufsize, size_t* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_create_int64(napi_env env, int64_t value, napi_value* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint,
                                 napi_value* result);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_is_array(napi_env env, napi_value value, bool* result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool* result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type,
                                     size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer,
                                   size_t byte_offset, napi_value* result);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_range_error(napi_env env, const char* code, const char* msg);

#define NAPI_AUTO_LENGTH SIZE_MAX

#ifdef __cplusplus
}
#endif
#endif /* NODE_API_MIN_H_ */
