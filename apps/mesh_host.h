// Host-side mesh container of the re-hosted C++ drivers: the pieces of the reference's Mesh class
// (src/lib/mesh.{h,cc}) that the drivers use outside the hot path -- OBJ I/O, Normalize,
// ApplyTransform, ReflectionSymmetrize -- in FP64 like the reference (FT = double, src/lib/types.h).
// The distance field and the solver run on the GPU behind the C-ABI (include/meshode_b200.h).
#pragma once
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace mo_app {

struct Mesh {
  std::vector<double> V;   // [nV,3]
  std::vector<int> F;      // [nF,3]
  double scale = 1.0;
  double pos[3] = {0.0, 0.0, 0.0};
  int nV() const { return (int)(V.size() / 3); }
  int nF() const { return (int)(F.size() / 3); }

  // src/lib/mesh.cc:14-42: lines of at most 255 characters, "v x y z" and "f a b c" records only,
  // "a/b/c" -> the leading index, 1-based, digits only (no negative indices).
  bool ReadOBJ(const char* filename) {
    std::ifstream is(filename);
    if (!is) return false;
    char buffer[256];
    while (is.getline(buffer, 256)) {
      std::istringstream str(buffer);
      std::string tag;
      str >> tag;
      if (tag == "v") {
        double x = 0, y = 0, z = 0;
        str >> x >> y >> z;
        V.push_back(x); V.push_back(y); V.push_back(z);
      } else if (tag == "f") {
        for (int j = 0; j < 3; ++j) {
          std::string tok;
          str >> tok;
          int id = 0;
          size_t p = 0;
          while (p < tok.size() && tok[p] != '/') { id = id * 10 + (tok[p] - '0'); ++p; }
          F.push_back(id - 1);
        }
      }
    }
    return true;
  }

  // src/lib/mesh.cc:44-64 (denormalised unless `normalized`; default stream formatting: 6 significant digits)
  bool WriteOBJ(const char* filename, bool normalized = false) const {
    std::ofstream os(filename);
    if (!os) return false;
    for (int i = 0; i < nV(); ++i) {
      double v[3];
      for (int j = 0; j < 3; ++j) v[j] = normalized ? V[3 * i + j] : V[3 * i + j] * scale + pos[j];
      os << "v " << v[0] << " " << v[1] << " " << v[2] << "\n";
    }
    for (int i = 0; i < nF(); ++i) os << "f " << F[3 * i] + 1 << " " << F[3 * i + 1] + 1 << " " << F[3 * i + 2] + 1 << "\n";
    return true;
  }

  // src/lib/mesh.cc:66-85 (the second min/max pass :86-95 is dead code)
  void Normalize() {
    double mn[3], mx[3];
    for (int j = 0; j < 3; ++j) {
      mn[j] = 1e30; mx[j] = -1e30;
      for (int i = 0; i < nV(); ++i) {
        if (V[3 * i + j] < mn[j]) mn[j] = V[3 * i + j];
        if (V[3 * i + j] > mx[j]) mx[j] = V[3 * i + j];
      }
    }
    const double e12 = (mx[1] - mn[1]) < (mx[2] - mn[2]) ? (mx[2] - mn[2]) : (mx[1] - mn[1]);
    const double e = (mx[0] - mn[0]) < e12 ? e12 : (mx[0] - mn[0]);
    scale = e * 1.1;
    for (int j = 0; j < 3; ++j) pos[j] = mn[j] - 0.05 * scale;
    for (int i = 0; i < nV(); ++i)
      for (int j = 0; j < 3; ++j) V[3 * i + j] = (V[3 * i + j] - pos[j]) / scale;
  }

  // src/lib/mesh.cc:98-105
  void ApplyTransform(const Mesh& m) {
    scale = m.scale;
    for (int j = 0; j < 3; ++j) pos[j] = m.pos[j];
    for (int i = 0; i < nV(); ++i)
      for (int j = 0; j < 3; ++j) V[3 * i + j] = (V[3 * i + j] - pos[j]) / scale;
  }

  // src/lib/mesh.cc:234-246
  void ReflectionSymmetrize() {
    const int vn = nV(), fn = nF();
    for (int i = 0; i < vn; ++i) { V.push_back(-V[3 * i]); V.push_back(V[3 * i + 1]); V.push_back(V[3 * i + 2]); }
    for (int i = 0; i < fn; ++i)
      for (int j = 0; j < 3; ++j) F.push_back(F[3 * i + j] + vn);
  }
};

}  // namespace mo_app
