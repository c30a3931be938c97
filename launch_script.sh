#!/bin/sh
# Launch template (same check-list as the reference's launch_script.sh):
#  (1) num-process: number of worker GPUs          (2) ips: host:ranks[,host:ranks...]
#  (3) master: rank-0 host                         (4) exec-file: training script
#  (5) socket_port: base port for rendezvous names (6) entry_point: 6 detect+profile, 7 profile, -1 none
#  (7) logical_graph / (8) strategy_file: intermediate files (written when entry_point is 6/7)
#  (9) parallel_degree: number of parallel transmissions (trees)   (10) profile_freq: re-profile period
# --mpi-path / --net-device are accepted for compatibility and ignored (torchrun, NVLink data path).
python -m adapcc_b200.launcher \
    --num-process 8 \
    --ips 127.0.0.1:8 \
    --master 127.0.0.1 \
    --exec-file train_ddp.py \
    --socket_port 5000 \
    --entry_point 7 \
    --logical_graph ./topology/logical_graph_8.xml \
    --strategy_file ./strategy/8.xml \
    --parallel_degree 4 \
    --profile_freq 500
