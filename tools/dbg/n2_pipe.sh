N=2
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) bench.py --gpus $N --value-only $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e6,1), round(d['ms_per_step'],3))"; }
export URNN_V2_PIPE=2 URNN_V2_GRID_E=147 URNN_V2_GRID_D=147
run "pipe+147 weak" ""
run "pipe+147 strong" "--scaling strong --height 4096 --width 4096 --steps 20 --warmup 3"
export URNN_V2_GRID_E=146 URNN_V2_GRID_D=146
run "pipe+146 weak" ""
run "pipe+146 strong" "--scaling strong --height 4096 --width 4096 --steps 20 --warmup 3"
