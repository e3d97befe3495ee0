#include "../stub_dealii.h"
