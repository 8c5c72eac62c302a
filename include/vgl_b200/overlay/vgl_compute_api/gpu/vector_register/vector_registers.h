// vector_registers.h of the B200 backend — per-lane scratch for user lambdas.
// Replaces vgl_compute_api/gpu/vector_register/vector_registers.h:3-75 (same macro names, same meaning): a "vector register"
// is an array of VECTOR_LENGTH values in device-accessible (managed) memory that a lambda indexes with its `vector_index`
// (= the lane id, 0..31, on this backend); the host folds it with register_*_reduce after the operator has returned
// (every operator of this backend ends with a stream synchronisation, so the host sees the final values).
#pragma once

#define VGLB_VEC_REGISTER(type, name, value)                         \
    type *reg_##name;                                                \
    MemoryAPI::allocate_array(&reg_##name, VECTOR_LENGTH);           \
    for (int vglb_i__ = 0; vglb_i__ < VECTOR_LENGTH; vglb_i__++) reg_##name[vglb_i__] = value;

#define VEC_REGISTER_INT(name, value) VGLB_VEC_REGISTER(int, name, value)
#define VEC_REGISTER_FLT(name, value) VGLB_VEC_REGISTER(float, name, value)
#define VEC_REGISTER_DBL(name, value) VGLB_VEC_REGISTER(double, name, value)

template <typename _T>
_T register_sum_reduce(_T *reg_name)
{
    _T sum = 0;
    for (int i = 0; i < VECTOR_LENGTH; i++) sum += reg_name[i];
    return sum;
}

template <typename _T>
_T register_max_reduce(_T *reg_name)
{
    _T result = reg_name[0];
    for (int i = 1; i < VECTOR_LENGTH; i++)
        if (reg_name[i] > result) result = reg_name[i];
    return result;
}

template <typename _T>
_T register_min_reduce(_T *reg_name)
{
    _T result = reg_name[0];
    for (int i = 1; i < VECTOR_LENGTH; i++)
        if (reg_name[i] < result) result = reg_name[i];
    return result;
}

template <typename _T>
void register_free(_T *reg_name)
{
    MemoryAPI::free_array(reg_name);
}
