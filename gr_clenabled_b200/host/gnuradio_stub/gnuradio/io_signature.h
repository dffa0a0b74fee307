#include <gnuradio/block.h>
