// empty stand-in: the reference includes it but uses nothing from it (dsp_dynamic.h:29)
