set -x
nvidia-smi -L
cd /root/repo
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -40
