// fb_veg.cu — mesh ingest and the displacement hand-off to rendering (N1).
//
// N2: reader for the Vega .veg text format with the reference loader's rules (VolumetricMesh(char*),
// src/3rdparty/vegafem/volumetricMesh/volumetricMesh.cpp:45-535; line handling volumetricMeshParser.cpp:132-211):
//   * lines starting with '#', CR or LF are skipped; *VERTICES then *ELEMENTS (TET) in that order;
//   * a vertex / element line is "index" followed by values, each found by seeking the next ' ' or ',' and scanning
//     from the character after it (this literal rule — including what it does with doubled separators — is kept);
//   * element vertices are 1-indexed in the file;
//   * *MATERIAL name / "ENU, density, E, nu"; *SET name / comma-separated 1-indexed elements on the following lines;
//     *REGION / "set, material"; regions are applied in file order (later ones win); set 0 is "allElements";
//   * elements no region covers get the LAST material, or — if the file has none — the reference's default material,
//     whose constructor arguments are swapped in the reference (ENuMaterial(name, density, E, nu) is called with
//     (E_default, nu_default, density_default)), giving density 1e9, E 0.45, nu 1000 (volumetricMesh.cpp:518, :40-42).
// *INCLUDE, CUBIC elements and MOONEYRIVLIN materials are rejected with FB_ERR_NOT_SUPPORTED (CorotationalLinearFEM
// itself throws on non-ENU materials, corotationalLinearFEM.cpp:61-64).
//
// N1: GPUPoly::applyFemDisplacements (src/implicit/OclPolygonizer.cpp:1543-1584) converts the displacement doubles to
// homogeneous float4 on the host, uploads them and runs ApplyVertexDeformations (data/opencl/Polygonizer.cl:1417-1427),
// out = restPos + deformation in float.  k_export_float4 does the conversion and the add on the device from the
// resident q, so the per-frame D2H + repack + H2D round trip disappears.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "fb_internal.h"

namespace {

void strip_spaces(char *s) {  // VolumetricMeshParser::removeWhitespace with numRetainedSpaces = 0
  char *w = s;
  for (char *p = s; *p; p++)
    if (*p != ' ') *w++ = *p;
  *w = 0;
}

bool next_line(FILE *f, char *buf, size_t n) {
  do {
    if (!fgets(buf, (int)n, f)) return false;
  } while (buf[0] == '#' || buf[0] == 13 || buf[0] == 10);
  size_t len = strlen(buf);
  if (len && buf[len - 1] == '\n') buf[len - 1] = 0;
  return true;
}

struct VegMesh {
  std::vector<double> verts;
  std::vector<int> tets;
  std::vector<double> E, nu, rho;
};

int parse_veg(const char *path, VegMesh &m) {
  FILE *f = fopen(path, "r");
  if (!f) { fb_set_error("could not open file %s", path); return FB_ERR_INVALID_ARGUMENT; }
  char line[4096];
  int numVertices = -1, numElements = -1, state = 0;
  int numMaterials = 0;
  // ---- first pass: vertices and elements
  while (next_line(f, line, sizeof(line))) {
    if (!strncmp(line, "*INCLUDE", 8)) { fclose(f); fb_set_error("%s: *INCLUDE is not supported", path); return FB_ERR_NOT_SUPPORTED; }
    if (state == 0 && !strncmp(line, "*VERTICES", 9)) {
      state = 1;
      if (!next_line(f, line, sizeof(line)) || sscanf(line, "%d", &numVertices) != 1 || numVertices < 0) {
        fclose(f); fb_set_error("file %s is not in the .veg format", path); return FB_ERR_BAD_MESH;
      }
      m.verts.reserve(3 * (size_t)numVertices);
      continue;
    }
    if (state == 1 && !strncmp(line, "*ELEMENTS", 9)) {
      state = 2;
      if (!next_line(f, line, sizeof(line))) { fclose(f); fb_set_error("file %s is not in the .veg format", path); return FB_ERR_BAD_MESH; }
      strip_spaces(line);
      if (strncmp(line, "TET", 3)) {
        fclose(f);
        fb_set_error("%s: element type %s is not on the path (TET only)", path, line);
        return FB_ERR_NOT_SUPPORTED;
      }
      if (!next_line(f, line, sizeof(line)) || sscanf(line, "%d", &numElements) != 1 || numElements < 0) {
        fclose(f); fb_set_error("file %s is not in the .veg format", path); return FB_ERR_BAD_MESH;
      }
      m.tets.reserve(4 * (size_t)numElements);
      continue;
    }
    if (state == 2 && line[0] == '*') state = 3;
    if (state == 1 || state == 2) {
      const int want = state == 1 ? 3 : 4;
      if ((state == 1 && (int)(m.verts.size() / 3) >= numVertices) || (state == 2 && (int)(m.tets.size() / 4) >= numElements)) {
        fclose(f); fb_set_error("mismatch in the number of %s in %s", state == 1 ? "vertices" : "elements", path); return FB_ERR_BAD_MESH;
      }
      char *ch = line;
      for (int i = 0; i < want; i++) {
        while (*ch != ' ' && *ch != ',' && *ch != 0) ch++;
        if (*ch == 0) { fclose(f); fb_set_error("error parsing line '%s' in file %s", line, path); return FB_ERR_BAD_MESH; }
        ch++;
        if (state == 1) {
          double v = 0.0;
          sscanf(ch, "%lf", &v);
          m.verts.push_back(v);
        } else {
          int v = 0;
          sscanf(ch, "%d", &v);
          m.tets.push_back(v - 1);
        }
      }
    }
    if (!strncmp(line, "*MATERIAL", 9)) numMaterials++;
  }
  if (numElements < 0 || numVertices < 0 || (int)(m.verts.size() / 3) != numVertices || (int)(m.tets.size() / 4) != numElements) {
    fclose(f); fb_set_error("incorrect number of vertices/elements: %s may not be in the .veg format", path); return FB_ERR_BAD_MESH;
  }
  // ---- second pass: materials, sets, regions
  rewind(f);
  struct Mat { std::string name; double density, E, nu; };
  std::vector<Mat> mats;
  std::vector<std::string> setNames(1, "allElements");
  std::vector<std::vector<int> > sets(1);
  for (int el = 0; el < numElements; el++) sets[0].push_back(el + 1);
  std::vector<std::pair<int, int> > regions;  // (material, set)
  state = 0;
  while (next_line(f, line, sizeof(line))) {
    if (state == 11 && line[0] == '*') state = 0;
    if (state == 0 && !strncmp(line, "*MATERIAL", 9)) {
      strip_spaces(line);
      Mat mt;
      mt.name = line + 9;
      if (!next_line(f, line, sizeof(line))) { fclose(f); fb_set_error("incorrect material in file %s", path); return FB_ERR_BAD_MESH; }
      strip_spaces(line);
      if (strncmp(line, "ENU", 3)) {
        fclose(f);
        fb_set_error("%s: material type '%s' is not supported (ENU only, as CorotationalLinearFEM requires)", path, line);
        return FB_ERR_NOT_SUPPORTED;
      }
      char *ch = line;
      while (*ch != ',' && *ch != 0) ch++;
      if (*ch == 0 || sscanf(ch + 1, "%lf,%lf,%lf", &mt.density, &mt.E, &mt.nu) != 3 ||
          !(mt.E > 0 && mt.nu > -1.0 && mt.nu < 0.5 && mt.density > 0)) {
        fclose(f); fb_set_error("incorrect material specification in file %s: %s", path, line); return FB_ERR_BAD_MESH;
      }
      mats.push_back(mt);
    }
    if (state == 0 && !strncmp(line, "*REGION", 7)) {
      if (!next_line(f, line, sizeof(line))) { fclose(f); fb_set_error("file %s is not in the .veg format", path); return FB_ERR_BAD_MESH; }
      strip_spaces(line);
      char *ch = line;
      while (*ch != ',' && *ch != 0) ch++;
      if (*ch == 0) { fclose(f); fb_set_error("error parsing region '%s' in %s", line, path); return FB_ERR_BAD_MESH; }
      *ch = 0;
      const std::string setName = line, matName = ch + 1;
      int si = -1, mi = -1;
      for (size_t k = 0; k < setNames.size(); k++) if (setNames[k] == setName) { si = (int)k; break; }
      for (size_t k = 0; k < mats.size(); k++) if (mats[k].name == matName) { mi = (int)k; break; }
      if (si < 0 || mi < 0) {
        fclose(f);
        fb_set_error("%s: %s name %s not found", path, si < 0 ? "set" : "material", si < 0 ? setName.c_str() : matName.c_str());
        return FB_ERR_BAD_MESH;
      }
      regions.push_back(std::make_pair(mi, si));
    }
    if (state == 0 && !strncmp(line, "*SET", 4)) {
      strip_spaces(line);
      setNames.push_back(line + 4);
      sets.push_back(std::vector<int>());
      state = 11;
      continue;  // the *SET line itself carries no elements (strtok on it yields a non-digit token in the reference)
    }
    if (state == 11) {
      strip_spaces(line);
      for (char *pch = strtok(line, ","); pch && isdigit((unsigned char)*pch); pch = strtok(NULL, ",")) sets.back().push_back(atoi(pch));
    }
  }
  fclose(f);
  // ---- materials to elements: regions in file order, then the default rule
  std::vector<int> elMat((size_t)numElements, -1);
  for (size_t rgn = 0; rgn < regions.size(); rgn++) {
    const std::vector<int> &s = sets[regions[rgn].second];
    for (size_t k = 0; k < s.size(); k++) {
      const int el = s[k] - 1;
      if (el < 0 || el >= numElements) { fb_set_error("%s: set element %d out of range", path, s[k]); return FB_ERR_BAD_MESH; }
      elMat[el] = regions[rgn].first;
    }
  }
  bool unassigned = false;
  for (int el = 0; el < numElements; el++) unassigned |= elMat[el] < 0;
  if (unassigned) {
    if (mats.empty()) {
      Mat d;
      d.name = "defaultMaterial";
      d.density = 1E9; d.E = 0.45; d.nu = 1000;  // the reference's swapped default (see the header comment)
      mats.push_back(d);
    }
    for (int el = 0; el < numElements; el++)
      if (elMat[el] < 0) elMat[el] = (int)mats.size() - 1;
  }
  m.E.resize((size_t)numElements); m.nu.resize((size_t)numElements); m.rho.resize((size_t)numElements);
  for (int el = 0; el < numElements; el++) {
    m.E[el] = mats[elMat[el]].E; m.nu[el] = mats[elMat[el]].nu; m.rho[el] = mats[elMat[el]].density;
  }
  return FB_OK;
}

template <typename T>
T *dup_array(const std::vector<T> &v) {
  T *p = (T *)malloc(sizeof(T) * (v.size() ? v.size() : 1));
  if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
  return p;
}

// out[i] = rest[i] + (float4)(q[3i], q[3i+1], q[3i+2], 0)   (float arithmetic, as ApplyVertexDeformations)
__global__ void k_export_float4(int n, const double *__restrict__ q, const double *__restrict__ x0, const float4 *__restrict__ rest,
                                float4 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 rp;
  if (rest) rp = rest[i];
  else rp = make_float4((float)x0[3 * (size_t)i], (float)x0[3 * (size_t)i + 1], (float)x0[3 * (size_t)i + 2], 1.0f);
  const float4 dp = make_float4((float)q[3 * (size_t)i], (float)q[3 * (size_t)i + 1], (float)q[3 * (size_t)i + 2], 0.0f);
  out[i] = make_float4(rp.x + dp.x, rp.y + dp.y, rp.z + dp.z, rp.w + dp.w);
}

}  // namespace

extern "C" {

int fb_veg_load(const char *path, int *num_vertices, int *num_tets, double **vertices, int **tets, double **E, double **nu,
                double **density) {
  if (!path || !num_vertices || !num_tets) { fb_set_error("bad arguments to fb_veg_load"); return FB_ERR_INVALID_ARGUMENT; }
  VegMesh m;
  FB_TRY(parse_veg(path, m));
  *num_vertices = (int)(m.verts.size() / 3);
  *num_tets = (int)(m.tets.size() / 4);
  if (vertices) *vertices = dup_array(m.verts);
  if (tets) *tets = dup_array(m.tets);
  if (E) *E = dup_array(m.E);
  if (nu) *nu = dup_array(m.nu);
  if (density) *density = dup_array(m.rho);
  return FB_OK;
}

void fb_veg_free(void *p) { free(p); }

int fb_veg_save(const char *path, int style, int nV, const double *vertices, int nT, const int *tets, const double *E,
                const double *nu, const double *density) {
  if (!path || nV < 0 || nT < 0 || (nV > 0 && !vertices) || (nT > 0 && !tets) || (style != FB_VEG_STYLE_FEMBRAIN && style != FB_VEG_STYLE_VEGA)) {
    fb_set_error("bad arguments to fb_veg_save");
    return FB_ERR_INVALID_ARGUMENT;
  }
  const bool haveMat = E && nu && density;
  if (style == FB_VEG_STYLE_VEGA && !haveMat && (E || nu || density)) {
    fb_set_error("fb_veg_save: E, nu and density must be given together");
    return FB_ERR_INVALID_ARGUMENT;
  }
  for (size_t i = 0; i < 4 * (size_t)nT; i++)
    if (tets[i] < 0 || tets[i] >= nV) { fb_set_error("fb_veg_save: tetrahedron %zu references a vertex outside [0, %d)", i / 4, nV); return FB_ERR_BAD_MESH; }
  // VolMeshIO::writeVega refuses an empty mesh (VolMeshIO.cpp:175-176)
  if (style == FB_VEG_STYLE_FEMBRAIN && (nV == 0 || nT == 0)) { fb_set_error("fb_veg_save: empty mesh"); return FB_ERR_BAD_MESH; }
  FILE *f = fopen(path, "w");
  if (!f) { fb_set_error("fb_veg_save: cannot write %s", path); return FB_ERR_INVALID_ARGUMENT; }
  if (style == FB_VEG_STYLE_FEMBRAIN) {
    fprintf(f, "# Vega Mesh File, Generated by FemBrain.\n# %d vertices, %d elements\n\n*VERTICES\n%d 3 0 0\n", nV, nT, nV);
    // operator<<(double) with the stream defaults = %g
    for (int i = 0; i < nV; i++) fprintf(f, "%d %g %g %g\n", i + 1, vertices[3 * (size_t)i], vertices[3 * (size_t)i + 1], vertices[3 * (size_t)i + 2]);
    fprintf(f, "\n*ELEMENTS\nTET\n%d 4 0\n", nT);
    for (int i = 0; i < nT; i++) {
      const int *t = tets + 4 * (size_t)i;
      fprintf(f, "%d %d %d %d %d\n", i + 1, t[0] + 1, t[1] + 1, t[2] + 1, t[3] + 1);
    }
    fprintf(f, "\n*MATERIAL BODY\nENU, 1000, 10000000, 0.45\n\n*REGION\nallElements, BODY\n");
  } else {
    fprintf(f, "# Vega mesh file.\n# %d vertices, %d elements\n\n*VERTICES\n%d 3 0 0\n", nV, nT, nV);
    for (int i = 0; i < nV; i++)
      fprintf(f, "%d %.15G %.15G %.15G\n", i + 1, vertices[3 * (size_t)i], vertices[3 * (size_t)i + 1], vertices[3 * (size_t)i + 2]);
    fprintf(f, "\n*ELEMENTS\nTET\n%d 4 0\n", nT);
    for (int i = 0; i < nT; i++) {
      const int *t = tets + 4 * (size_t)i;
      fprintf(f, "%d %d %d %d %d\n", i + 1, t[0] + 1, t[1] + 1, t[2] + 1, t[3] + 1);
    }
    fprintf(f, "\n");
    if (haveMat && nT > 0) {
      // distinct (density, E, nu) triples in order of first appearance; element -> material
      std::vector<int> first;           // element that introduced material k
      std::vector<int> mat((size_t)nT);
      for (int el = 0; el < nT; el++) {
        int k = 0;
        for (; k < (int)first.size(); k++) {
          const int e0 = first[(size_t)k];
          if (density[e0] == density[el] && E[e0] == E[el] && nu[e0] == nu[el]) break;
        }
        if (k == (int)first.size()) first.push_back(el);
        mat[(size_t)el] = k;
      }
      const int nM = (int)first.size();
      for (int k = 0; k < nM; k++) {
        const int e0 = first[(size_t)k];
        fprintf(f, "*MATERIAL material_%d\nENU, %.15G, %.15G, %.15G\n\n", k, density[e0], E[e0], nu[e0]);
      }
      if (nM > 1) {
        for (int k = 0; k < nM; k++) {
          fprintf(f, "*SET set_%d\n", k);
          int count = 0;
          for (int el = 0; el < nT; el++) {
            if (mat[(size_t)el] != k) continue;
            fprintf(f, "%d, ", el + 1);
            if (++count == 8) { fprintf(f, "\n"); count = 0; }
          }
          if (count != 0) fprintf(f, "\n");
          fprintf(f, "\n");
        }
        for (int k = 0; k < nM; k++) fprintf(f, "*REGION\nset_%d, material_%d\n\n", k, k);
      } else {
        fprintf(f, "*REGION\nallElements, material_0\n\n");
      }
    }
  }
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) { fb_set_error("fb_veg_save: write to %s failed", path); return FB_ERR_INVALID_ARGUMENT; }
  return FB_OK;
}

// TetMesh(char* filename, int specialFileType = 0), src/3rdparty/vegafem/volumetricMesh/tetMesh.cpp:45-127: TetGen's
// <base>.node ("numVertices 3", then "index x y z", 1-indexed and consecutive) and <base>.ele ("numElements 4", then
// "index v0 v1 v2 v3", vertices 1-indexed); comment and blank lines skipped as everywhere in the parser; one material
// for all elements, E = 1e8, nu = 0.45, density = 1000 (:47-49, setSingleMaterial :126).  The reference throws an int
// on every irregularity (2-6); here: FB_ERR_INVALID_ARGUMENT for a file that cannot be opened, FB_ERR_BAD_MESH otherwise.
// (That constructor cannot serve as the checker: its setSingleMaterial writes elementMaterial[] through a pointer the
// constructor never allocates, volumetricMesh.cpp:1002-1034, and crashes when compiled here — parity unpinned.)
int fb_tetgen_load(const char *basename, int *num_vertices, int *num_tets, double **vertices, int **tets, double **E, double **nu,
                   double **density) {
  if (!basename || !num_vertices || !num_tets) { fb_set_error("bad arguments to fb_tetgen_load"); return FB_ERR_INVALID_ARGUMENT; }
  char line[4096];
  std::string path = std::string(basename) + ".node";
  FILE *f = fopen(path.c_str(), "r");
  if (!f) { fb_set_error("could not open file %s", path.c_str()); return FB_ERR_INVALID_ARGUMENT; }
  int nV = 0, nT = 0, dim = 0;
  if (!next_line(f, line, sizeof(line)) || sscanf(line, "%d %d", &nV, &dim) != 2 || dim != 3 || nV < 0) {
    fclose(f); fb_set_error("%s: not a 3D TetGen node file", path.c_str()); return FB_ERR_BAD_MESH;
  }
  std::vector<double> v(3 * (size_t)nV);
  for (int i = 0; i < nV; i++) {
    int index = 0;
    if (!next_line(f, line, sizeof(line)) || sscanf(line, "%d %lf %lf %lf", &index, &v[3 * (size_t)i], &v[3 * (size_t)i + 1], &v[3 * (size_t)i + 2]) != 4 ||
        index != i + 1) {
      fclose(f); fb_set_error("%s: vertex line %d is missing, malformed or out of sequence", path.c_str(), i + 1); return FB_ERR_BAD_MESH;
    }
  }
  fclose(f);
  path = std::string(basename) + ".ele";
  f = fopen(path.c_str(), "r");
  if (!f) { fb_set_error("could not open file %s", path.c_str()); return FB_ERR_INVALID_ARGUMENT; }
  if (!next_line(f, line, sizeof(line)) || sscanf(line, "%d %d", &nT, &dim) != 2 || nT < 0) {
    fclose(f); fb_set_error("%s: not a TetGen element file", path.c_str()); return FB_ERR_BAD_MESH;
  }
  if (dim != 4) { fclose(f); fb_set_error("%s: not a tet mesh file (%d vertices per tet encountered)", path.c_str(), dim); return FB_ERR_BAD_MESH; }
  std::vector<int> t(4 * (size_t)nT);
  for (int i = 0; i < nT; i++) {
    int index = 0, w[4];
    if (!next_line(f, line, sizeof(line)) || sscanf(line, "%d %d %d %d %d", &index, &w[0], &w[1], &w[2], &w[3]) != 5 || index != i + 1) {
      fclose(f); fb_set_error("%s: element line %d is missing, malformed or out of sequence", path.c_str(), i + 1); return FB_ERR_BAD_MESH;
    }
    for (int j = 0; j < 4; j++) t[4 * (size_t)i + j] = w[j] - 1;  // vertices are 1-indexed in .ele files
  }
  fclose(f);
  *num_vertices = nV;
  *num_tets = nT;
  if (vertices) *vertices = dup_array(v);
  if (tets) *tets = dup_array(t);
  if (E) *E = dup_array(std::vector<double>((size_t)nT, 1E8));
  if (nu) *nu = dup_array(std::vector<double>((size_t)nT, 0.45));
  if (density) *density = dup_array(std::vector<double>((size_t)nT, 1000.0));
  return FB_OK;
}

int fb_create_from_veg(fb_context **out, const char *path, int num_fixed_vertices, const int *fixed_vertices, const fb_params *params) {
  if (!out || !path) { fb_set_error("bad arguments to fb_create_from_veg"); return FB_ERR_INVALID_ARGUMENT; }
  VegMesh m;
  FB_TRY(parse_veg(path, m));
  return fb_create_with_materials(out, (int)(m.verts.size() / 3), m.verts.data(), (int)(m.tets.size() / 4), m.tets.data(),
                                  num_fixed_vertices, fixed_vertices, m.E.data(), m.nu.data(), m.rho.data(), params);
}

int fb_export_positions_float4(fb_context *c, int count, const float *rest_xyzw, float *out_xyzw) {
  if (!c || count < 0 || count > c->nV || !out_xyzw) { fb_set_error("bad arguments to fb_export_positions_float4"); return FB_ERR_INVALID_ARGUMENT; }
  if (cudaSetDevice(c->device) != cudaSuccess) return FB_ERR_CUDA;
  if (count == 0) return FB_OK;
  float4 *dRest = nullptr, *dOut = nullptr;
  FB_CUDA(cudaMalloc(&dOut, sizeof(float4) * (size_t)count));
  if (rest_xyzw) {
    FB_CUDA(cudaMalloc(&dRest, sizeof(float4) * (size_t)count));
    FB_CUDA(cudaMemcpyAsync(dRest, rest_xyzw, sizeof(float4) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
  }
  k_export_float4<<<(count + 255) / 256, 256, 0, c->stream>>>(count, c->q, c->x0, dRest, dOut);
  c->launches++;
  FB_CUDA(cudaMemcpyAsync(out_xyzw, dOut, sizeof(float4) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  fb_dev_free(dRest);
  fb_dev_free(dOut);
  return FB_OK;
}

int fb_export_positions_float4_dev(fb_context *c, int count, const float *rest_xyzw_dev, float *out_xyzw_dev) {
  if (!c || count < 0 || count > c->nV || !out_xyzw_dev) { fb_set_error("bad arguments to fb_export_positions_float4_dev"); return FB_ERR_INVALID_ARGUMENT; }
  if (cudaSetDevice(c->device) != cudaSuccess) return FB_ERR_CUDA;
  if (count == 0) return FB_OK;
  k_export_float4<<<(count + 255) / 256, 256, 0, c->stream>>>(count, c->q, c->x0, (const float4 *)rest_xyzw_dev, (float4 *)out_xyzw_dev);
  c->launches++;
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

}  // extern "C"
