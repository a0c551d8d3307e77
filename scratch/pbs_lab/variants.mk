# variants compared by ./lab; the first one is the reference the others are checked against bit for bit
VARIANTS := r1 final fft_only fft_nosync mac_light
R1 := -DWS_MAC_REUSE=0 -DWS_REVMAP=0 -DWS_TWB_INV=0 -DWS_ROT_LATE=0 -DWS_MAC_REGS3=120
FLAGS_r1 := $(R1)
FLAGS_final :=
FLAGS_fft_only := $(R1) -DWS_DIAG=32
FLAGS_fft_nosync := $(R1) -DWS_DIAG=128
FLAGS_mac_light := $(R1) -DWS_DIAG=4
