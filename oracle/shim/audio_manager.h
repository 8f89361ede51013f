/*
 * oracle shim (test infrastructure): shadows /root/reference/include/audio_manager.h, which
 * resolution.c:2 includes but never uses and which would pull in OpenAL/libsndfile headers.
 */
#pragma once
