// Lookup / folding helpers over the flat (name, host pointer, shape) table a net is created from.
#pragma once
#include "common.cuh"
#include <math.h>
#include <string>
#include <string.h>
#include <vector>

namespace mimamo {

struct TensorTable {
  const mimamo_tensor_desc* t;
  int n;
  const mimamo_tensor_desc* find(const std::string& name) const {
    for (int i = 0; i < n; ++i)
      if (name == t[i].name) return &t[i];
    return nullptr;
  }
  // returns data pointer or nullptr (and sets the error) if missing / wrong element count
  const float* get(const std::string& name, int64_t numel) const {
    const mimamo_tensor_desc* d = find(name);
    if (!d) { set_error("state_dict is missing key '%s'", name.c_str()); return nullptr; }
    int64_t cnt = 1;
    for (int i = 0; i < d->ndim; ++i) cnt *= d->shape[i];
    if (cnt != numel) { set_error("tensor '%s' has %lld elements, expected %lld", name.c_str(), (long long)cnt, (long long)numel); return nullptr; }
    return d->data_host;
  }
};

// eval-mode BatchNorm as y = x*scale + shift (optionally absorbing a preceding bias)
inline bool fold_bn(const TensorTable& T, const std::string& bn, int C, float eps, const float* bias,
                    std::vector<float>& scale, std::vector<float>& shift) {
  const float* g = T.get(bn + ".weight", C);
  const float* b = T.get(bn + ".bias", C);
  const float* m = T.get(bn + ".running_mean", C);
  const float* v = T.get(bn + ".running_var", C);
  if (!g || !b || !m || !v) return false;
  scale.resize(C); shift.resize(C);
  for (int i = 0; i < C; ++i) {
    const double s = (double)g[i] / sqrt((double)v[i] + (double)eps);
    scale[i] = (float)s;
    shift[i] = (float)((double)b[i] + s * ((bias ? (double)bias[i] : 0.0) - (double)m[i]));
  }
  return true;
}

}  // namespace mimamo
