#include "../stub_all.h"
