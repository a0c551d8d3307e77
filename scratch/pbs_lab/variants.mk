VARIANTS := r1 final ws2
R1 := -DWS_MAC_REUSE=0 -DWS_REVMAP=0 -DWS_TWB_INV=0 -DWS_ROT_LATE=0 -DWS_MAC_REGS3=120 -DWS_TMEM_ST=0
FLAGS_r1 := $(R1)
FLAGS_final :=
SRC_ws2 := pbs_ws2_kernel.cu
FLAGS_ws2 := -DPBS_WS2_NS=ns_ws2 -DPBS_WS2_LAUNCH_NAME=launch_ws2
VARIANTS += ws2t
SRC_ws2t := pbs_ws2_kernel.cu
FLAGS_ws2t := -DPBS_WS2_NS=ns_ws2t -DPBS_WS2_LAUNCH_NAME=launch_ws2t -DWS2_TIMING=1
VARIANTS += ws2w
SRC_ws2w := pbs_ws2_kernel.cu
FLAGS_ws2w := -DPBS_WS2_NS=ns_ws2w -DPBS_WS2_LAUNCH_NAME=launch_ws2w -DWS2_ORDER=0x5761234
VARIANTS += ws2f
SRC_ws2f := pbs_ws2_kernel.cu
FLAGS_ws2f := -DPBS_WS2_NS=ns_ws2f -DPBS_WS2_LAUNCH_NAME=launch_ws2f -DWS2_ORDER=0x5712364
VARIANTS += ws2g
SRC_ws2g := pbs_ws2_kernel.cu
FLAGS_ws2g := -DPBS_WS2_NS=ns_ws2g -DPBS_WS2_LAUNCH_NAME=launch_ws2g -DWS2_ORDER=0x5716234
VARIANTS += ws2h
SRC_ws2h := pbs_ws2_kernel.cu
FLAGS_ws2h := -DPBS_WS2_NS=ns_ws2h -DPBS_WS2_LAUNCH_NAME=launch_ws2h -DWS2_ORDER=0x5176234
VARIANTS += ws2w112
SRC_ws2w112 := pbs_ws2_kernel.cu
FLAGS_ws2w112 := -DPBS_WS2_NS=ns_ws2w112 -DPBS_WS2_LAUNCH_NAME=launch_ws2w112 -DWS2_ORDER=0x5761234 -DWS2_MAC_REGS=112
VARIANTS += ws2w128
SRC_ws2w128 := pbs_ws2_kernel.cu
FLAGS_ws2w128 := -DPBS_WS2_NS=ns_ws2w128 -DPBS_WS2_LAUNCH_NAME=launch_ws2w128 -DWS2_ORDER=0x5761234 -DWS2_MAC_REGS=128
