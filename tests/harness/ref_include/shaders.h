// TEST ONLY: lets the oracle build of example_main.cpp take shaders.h from our host layer while
// geometry.h / tgaimage.h / our_gl.h come from the reference checkout (-I /root/reference).
#include "../../../tinyrenderder_b200/host/shaders.h"
