// client_internal.h — accessors of the opaque fsc_client for the other host-side translation units (not installed).
#pragma once
#include <stdint.h>
#include <string>

#include "../../include/fhe_sign_cuda.h"

fsc_params fsc_client_params(const fsc_client* c);
fsc_noise_params fsc_client_noise(const fsc_client* c);
uint64_t fsc_client_seed(const fsc_client* c);
uint64_t fsc_client_enc_counter(const fsc_client* c);
void fsc_client_set_enc_counter(fsc_client* c, uint64_t v);
void fsc_client_set_error(const std::string& msg);      // message returned by fsc_client_last_error(NULL) on this thread
