// placeholder replaced below by the cuadmm_exe front end
#include <stdio.h>
#include "../../include/cuadmm_b200.h"
int main(int argc, char** argv) { printf("%s\n", cuadmm_version()); return 0; }
