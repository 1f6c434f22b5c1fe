// Stub for glad -- TEST INFRASTRUCTURE. Only what src/utils.h:35-55 references.
#pragma once
typedef unsigned int GLenum;
typedef unsigned int GLuint;
#define GL_NO_ERROR 0
#define GL_INVALID_ENUM 0x0500
#define GL_INVALID_VALUE 0x0501
#define GL_INVALID_OPERATION 0x0502
#define GL_OUT_OF_MEMORY 0x0505
#define GL_INVALID_FRAMEBUFFER_OPERATION 0x0506
inline GLenum glGetError() { return GL_NO_ERROR; }
