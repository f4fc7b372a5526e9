// Marker translation unit: enables DDPT_HAVE_LOADER (see ops/build.py); the binding itself lives in
// data/host_loader.cpp so that the loader has no torch dependency.
