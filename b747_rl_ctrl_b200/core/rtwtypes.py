"""Simulink rtwtypes aliases (mirror of core/rtwtypes.py; only real_T is used, core/model.py:16-18)."""
import ctypes

real_T = ctypes.c_double       # core/rtwtypes.py:27
real64_T = ctypes.c_double
real32_T = ctypes.c_float
int32_T = ctypes.c_int32
uint32_T = ctypes.c_uint32
int8_T = ctypes.c_int8
uint8_T = ctypes.c_uint8
boolean_T = ctypes.c_uint8
time_T = ctypes.c_double
