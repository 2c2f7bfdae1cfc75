/* ORACLE — TEST INFRASTRUCTURE ONLY. Thin driver over the REFERENCE's own tangent-space generator: compiled together
 * with /root/reference/Extern/MikkTSpace/mikktspace.c (where it lies; never copied) into oracle/_ref/libmikktspace_ref.so
 * by oracle/Makefile's `ref` target. It feeds the library the way Include/Althea/GeometryUtilities.h:51-155 does
 * (three vertices per face, basic callback only) and is what tests/test_tangent_space.py and
 * tests/golden/make_tangent_golden.py hold althea_host_compute_tangent_space to. */
#include "mikktspace.h"

#include <stdint.h>

typedef struct {
  const float* position;
  const float* normal;
  const float* uv;
  int faces;
  float* tangent;
  float* sign;
} Soup;

static int numFaces(const SMikkTSpaceContext* c) { return ((Soup*)c->m_pUserData)->faces; }
static int vertsOfFace(const SMikkTSpaceContext* c, const int f) { return f < ((Soup*)c->m_pUserData)->faces ? 3 : 0; }
static void getPosition(const SMikkTSpaceContext* c, float out[], const int f, const int v) {
  const float* p = ((Soup*)c->m_pUserData)->position + 3 * (3 * f + v);
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}
static void getNormal(const SMikkTSpaceContext* c, float out[], const int f, const int v) {
  const float* p = ((Soup*)c->m_pUserData)->normal + 3 * (3 * f + v);
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}
static void getTexCoord(const SMikkTSpaceContext* c, float out[], const int f, const int v) {
  const float* p = ((Soup*)c->m_pUserData)->uv + 2 * (3 * f + v);
  out[0] = p[0]; out[1] = p[1];
}
static void setBasic(const SMikkTSpaceContext* c, const float t[], const float s, const int f, const int v) {
  Soup* m = (Soup*)c->m_pUserData;
  float* o = m->tangent + 3 * (3 * f + v);
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
  m->sign[3 * f + v] = s;
}

int ref_mikktspace(const float* position, const float* normal, const float* uv, int faces, float* tangent, float* sign) {
  Soup soup = {position, normal, uv, faces, tangent, sign};
  SMikkTSpaceInterface iface = {0};
  SMikkTSpaceContext ctx = {0};
  iface.m_getNumFaces = numFaces;
  iface.m_getNumVerticesOfFace = vertsOfFace;
  iface.m_getPosition = getPosition;
  iface.m_getNormal = getNormal;
  iface.m_getTexCoord = getTexCoord;
  iface.m_setTSpaceBasic = setBasic;
  ctx.m_pInterface = &iface;
  ctx.m_pUserData = &soup;
  return genTangSpaceDefault(&ctx);
}
