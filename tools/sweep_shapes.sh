#!/bin/bash
# batch-size / shape sweep of the propagation step on one GPU (tool): one JSON line per shape
for b in 1 2 4 16 32; do echo "kitti K3 T18 fwdbwd B=$b"; python tools/quick_step.py --batch $b --steps 5; done
echo "kitti K7 T12 B=4"; python tools/quick_step.py --kernel 7 --iters 12 --batch 4 --steps 3
echo "kitti fwd B=16"; python tools/quick_step.py --mode fwd --batch 16 --steps 5
echo "kitti fwd B=1"; python tools/quick_step.py --mode fwd --batch 1 --steps 20
echo "nyu fwdbwd B=1"; python tools/quick_step.py --workload nyu --batch 1 --steps 20
echo "nyu fwdbwd B=4"; python tools/quick_step.py --workload nyu --batch 4 --steps 20
