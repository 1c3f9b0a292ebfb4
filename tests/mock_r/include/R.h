/* stand-in for R.h (tests/mock_r): TEST INFRASTRUCTURE ONLY */
#ifndef MOCK_R_H
#define MOCK_R_H
#include <stdlib.h>
#include <stdio.h>
#endif
