#ifndef MOCK_R_H
#define MOCK_R_H
#include "Rinternals.h"
#endif
