#!/bin/sh
# Launch template. Override with environment variables, append script arguments after "--":
#   NP=8 HOSTS=127.0.0.1:8 ENTRY=7 sh launch_script.sh -- --model vgg16 --steps 20
#
#   NP      worker GPUs in total              HOSTS   host:ranks[,host:ranks...] (first host = rank 0)
#   SCRIPT  training script                   PORT    base port for rendezvous names
#   ENTRY   6 detect + profile + synthesise, 7 profile + synthesise, -1 use STRATEGY as it is
#   GRAPH / STRATEGY   intermediate files (written when ENTRY is 6 or 7)
#   DEGREE  parallel transmissions (trees)    FREQ    steps between reconstruct_topology calls
# (--mpi-path / --net-device of the reference are accepted by the launcher and ignored: torchrun, NVLink data path)
: "${NP:=8}" "${HOSTS:=127.0.0.1:$NP}" "${SCRIPT:=train_ddp.py}" "${PORT:=5000}" "${ENTRY:=7}"
: "${GRAPH:=./topology/logical_graph_8.xml}" "${STRATEGY:=./strategy/8.xml}" "${DEGREE:=4}" "${FREQ:=500}"
[ "$1" = "--" ] && shift
exec python -m adapcc_b200.launcher --num-process "$NP" --ips "$HOSTS" --master "${HOSTS%%:*}" --exec-file "$SCRIPT" \
    --socket_port "$PORT" --entry_point "$ENTRY" --logical_graph "$GRAPH" --strategy_file "$STRATEGY" \
    --parallel_degree "$DEGREE" --profile_freq "$FREQ" -- "$@"
