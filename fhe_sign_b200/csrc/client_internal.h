// client_internal.h — accessors of the opaque fsc_client for the other host-side translation units (not installed).
#pragma once
#include <stdint.h>
#include <string>

#include "../../include/fhe_sign_cuda.h"

fsc_params fsc_client_params(const fsc_client* c);
fsc_noise_params fsc_client_noise(const fsc_client* c);
const uint8_t* fsc_client_master_key(const fsc_client* c);      // 32 bytes
// rebuilds secret + server keys from a persisted 256-bit master key; encryption randomness is drawn fresh from the OS
fsc_status fsc_client_keygen_from_master(const fsc_params* params, const fsc_noise_params* noise, const uint8_t* master32, fsc_client** out);
// range checks shared by keygen and the file loader (shift amounts derive from these fields)
bool fsc_params_plausible(const fsc_params& p, const fsc_noise_params& np, std::string* why);
void fsc_client_set_error(const std::string& msg);      // message returned by fsc_client_last_error(NULL) on this thread
