// TEST ONLY: CPU restatement of the post passes (main.cpp:269-362, 756-786) for the ORACLE build of
// tinyrenderder_b200/host/example_main.cpp; the device build calls gl_ssao()/gl_zbuffer_image()/gl_composite_ao().
#pragma once
#include <cstring>
#include <vector>
struct TrbCtx;
#define TRB_OK 0
#define TRB_E_ARG -1
static const double* orc_view_depth(TrbCtx*, int, int* w, int* h);
static const unsigned char* orc_view_color(TrbCtx*, int);
static int g_w, g_h;
static const unsigned char* g_color;
#define ORC_POST_PASSES_ONLY 1
#include "../../../oracle/post_restate.inc"
static const double* orc_view_depth(TrbCtx*, int, int* w, int* h) { *w = g_w; *h = g_h; return zbuffer.data(); }
static const unsigned char* orc_view_color(TrbCtx*, int) { return g_color; }
