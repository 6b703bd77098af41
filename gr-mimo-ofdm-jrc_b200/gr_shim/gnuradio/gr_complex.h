#include <gnuradio/shim_runtime.h>
