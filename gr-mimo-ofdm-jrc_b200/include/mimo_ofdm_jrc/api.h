// Export macro of the B200 radar-path blocks (same name as the reference's include/mimo_ofdm_jrc/api.h:28-32).
#pragma once
#include <gnuradio/attributes.h>
#ifdef gnuradio_mimo_ofdm_jrc_EXPORTS
#define MIMO_OFDM_JRC_API __GR_ATTR_EXPORT
#else
#define MIMO_OFDM_JRC_API __GR_ATTR_IMPORT
#endif
