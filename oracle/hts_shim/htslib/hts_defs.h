// htslib-API shim (oracle build only): see ../hts_shim.cpp
#pragma once
