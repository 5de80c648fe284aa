/*
 * gomath.h — Go standard library math.Sin/Cos/Tan/Asin/Acos/Atan/Atan2 restated in C
 * (see gomath.c).  TEST INFRASTRUCTURE ONLY.
 */
#ifndef NBODY_GOMATH_H
#define NBODY_GOMATH_H
#ifdef __cplusplus
extern "C" {
#endif
double go_sin(double x);
double go_cos(double x);
double go_tan(double x);
double go_atan(double x);
double go_asin(double x);
double go_acos(double x);
double go_atan2(double y, double x);
#ifdef __cplusplus
}
#endif
#endif
