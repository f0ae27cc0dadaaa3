// gl_shim.cpp -- no-op GL, and a GLUT main loop that just pumps the idle callback (see gl_shim.h).
#include "gl_shim.h"

#include <cstdio>
#include <cstdlib>

static void (*g_idle)(void) = nullptr;
static void (*g_display)(void) = nullptr;

extern "C" {
void glDisable(GLenum) {}
void glViewport(GLint, GLint, GLsizei, GLsizei) {}
void glMatrixMode(GLenum) {}
void glLoadIdentity(void) {}
void glOrtho(GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble) {}
void glPixelZoom(GLfloat, GLfloat) {}
void glRasterPos2f(GLfloat, GLfloat) {}
void glDrawPixels(GLsizei, GLsizei, GLenum, GLenum, const void*) {}

void glutInit(int*, char**) {}
void glutInitDisplayMode(unsigned int) {}
void glutInitWindowSize(int, int) {}
int glutCreateWindow(const char*) { return 1; }
void glutPositionWindow(int, int) {}
void glutMouseFunc(void (*)(int, int, int, int)) {}
void glutReshapeFunc(void (*)(int, int)) {}
void glutDisplayFunc(void (*f)(void)) { g_display = f; }
void glutKeyboardFunc(void (*)(unsigned char, int, int)) {}
void glutKeyboardUpFunc(void (*)(unsigned char, int, int)) {}
void glutIdleFunc(void (*f)(void)) { g_idle = f; }
void glutSpecialFunc(void (*)(int, int, int)) {}
void glutSpecialUpFunc(void (*)(int, int, int)) {}
void glutMotionFunc(void (*)(int, int)) {}
void glutSwapBuffers(void) {}

// TINSEL_HEADLESS_FRAMES bounds the loop for non-batch runs (each frame = 16 Render() calls,
// main.cpp:240-251); batch runs leave through exit() when the next numbered scene is missing.
void glutMainLoop(void)
{
    const char* lim = getenv("TINSEL_HEADLESS_FRAMES");
    long frames = lim ? atol(lim) : -1;
    void (*f)(void) = g_idle ? g_idle : g_display;
    while (f && frames != 0) {
        f();
        if (frames > 0) --frames;
    }
    exit(0);
}
}
