/* adds the steady-clock flag time_utils.cc asks for to the hand-written config.h of oracle/ref_config */
#pragma once
#include "../../../oracle/ref_config/config.h"
#define LIZARDFS_HAVE_STD_CHRONO_STEADY_CLOCK
