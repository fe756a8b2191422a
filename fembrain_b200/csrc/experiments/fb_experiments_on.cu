// fb_experiments_on.cu — marks a library built WITH csrc/experiments/ (python -m fembrain_b200.build --experiments).
extern "C" int fb_experiments_built(void) { return 1; }
