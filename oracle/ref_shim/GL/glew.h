// oracle/ref_shim/GL/glew.h -- TEST INFRASTRUCTURE ONLY (part of the oracle/_ref recipe).
//
// Stand-in for <GL/glew.h> so that the reference's own lib/Pangolin_IOWrapper/Keyframe.h can be
// compiled, unmodified and from where it lies under /root/reference, without an OpenGL stack.
// Every GL entry point that header calls is declared here; glBufferData records the vertex buffer
// Keyframe::computeVbo uploads (Keyframe.h:148-150) so that oracle/ref_keyframe.cpp can hand it back.
// The headers Keyframe.h forgets to include itself (memcpy, assert, rand, uint64_t) come from here too.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
typedef long GLsizeiptr;

#define GL_ARRAY_BUFFER 0x8892
#define GL_STATIC_DRAW 0x88E4
#define GL_FLOAT 0x1406
#define GL_UNSIGNED_BYTE 0x1401
#define GL_VERTEX_ARRAY 0x8074
#define GL_COLOR_ARRAY 0x8076
#define GL_POINTS 0x0000
#define GL_LINES 0x0001

namespace ref_shim {
struct GlCapture {
  std::vector<unsigned char> bufferData;  // bytes of the last glBufferData call
  int genCalls = 0, deleteCalls = 0;
};
inline GlCapture &capture() {
  static thread_local GlCapture c;
  return c;
}
}  // namespace ref_shim

inline void glGenBuffers(GLsizei, GLuint *b) { *b = (GLuint)++ref_shim::capture().genCalls; }
inline void glDeleteBuffers(GLsizei, const GLuint *) { ref_shim::capture().deleteCalls++; }
inline void glBindBuffer(GLenum, GLuint) {}
inline void glBufferData(GLenum, GLsizeiptr size, const void *data, GLenum) {
  auto &c = ref_shim::capture();
  c.bufferData.assign((const unsigned char *)data, (const unsigned char *)data + size);
}
inline void glPushMatrix() {}
inline void glPopMatrix() {}
inline void glMultMatrixf(const GLfloat *) {}
inline void glVertexPointer(GLint, GLenum, GLsizei, const void *) {}
inline void glColorPointer(GLint, GLenum, GLsizei, const void *) {}
inline void glEnableClientState(GLenum) {}
inline void glDisableClientState(GLenum) {}
inline void glDrawArrays(GLenum, GLint, GLsizei) {}
inline void glColor3f(GLfloat, GLfloat, GLfloat) {}
inline void glBegin(GLenum) {}
inline void glEnd() {}
inline void glVertex3f(GLfloat, GLfloat, GLfloat) {}
