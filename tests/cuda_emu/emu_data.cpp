// TEST INFRASTRUCTURE ONLY -- tacotron_b200/csrc/data.cu compiled for the host emulation (see emu.h)
#include "emu.h"
#include "../../tacotron_b200/csrc/data.cu"
