// gl_shim.h -- headless stand-in for the GL / GLUT symbols tinsel's src/main.cpp uses.
//
// main.cpp includes GL/GLUT headers only under _WIN32 / __APPLE__ (main.cpp:13-29) yet calls
// gl*/glut* unconditionally (main.cpp:285-302, 511-528), so on Linux it needs *some* declaration of
// them.  This header is force-included (-include) when compiling the UNMODIFIED main.cpp for the
// headless drop-in check; gl_shim.cpp implements the calls as no-ops and glutMainLoop() as a loop
// over the idle callback (tinsel's batch mode exits by itself, main.cpp:310-328 / :128-132).
#pragma once

typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
typedef double GLdouble;

#define GL_BLEND 0x0BE2
#define GL_LIGHTING 0x0B50
#define GL_DEPTH_TEST 0x0B71
#define GL_CULL_FACE 0x0B44
#define GL_PROJECTION 0x1701
#define GL_MODELVIEW 0x1700
#define GL_RGBA 0x1908
#define GL_FLOAT 0x1406

#define GLUT_RGBA 0
#define GLUT_DOUBLE 2
#define GLUT_DEPTH 16
#define GLUT_DOWN 0
#define GLUT_UP 1
#define GLUT_KEY_LEFT 100
#define GLUT_KEY_UP 101
#define GLUT_KEY_RIGHT 102
#define GLUT_KEY_DOWN 103

extern "C" {
void glDisable(GLenum cap);
void glViewport(GLint x, GLint y, GLsizei w, GLsizei h);
void glMatrixMode(GLenum mode);
void glLoadIdentity(void);
void glOrtho(GLdouble l, GLdouble r, GLdouble b, GLdouble t, GLdouble n, GLdouble f);
void glPixelZoom(GLfloat x, GLfloat y);
void glRasterPos2f(GLfloat x, GLfloat y);
void glDrawPixels(GLsizei w, GLsizei h, GLenum format, GLenum type, const void* data);

void glutInit(int* argc, char** argv);
void glutInitDisplayMode(unsigned int mode);
void glutInitWindowSize(int w, int h);
int glutCreateWindow(const char* title);
void glutPositionWindow(int x, int y);
void glutMouseFunc(void (*f)(int, int, int, int));
void glutReshapeFunc(void (*f)(int, int));
void glutDisplayFunc(void (*f)(void));
void glutKeyboardFunc(void (*f)(unsigned char, int, int));
void glutKeyboardUpFunc(void (*f)(unsigned char, int, int));
void glutIdleFunc(void (*f)(void));
void glutSpecialFunc(void (*f)(int, int, int));
void glutSpecialUpFunc(void (*f)(int, int, int));
void glutMotionFunc(void (*f)(int, int));
void glutSwapBuffers(void);
void glutMainLoop(void);
}
