cd /root/repo
timeout 600 python tools/_dbg1.py 2>&1 | tail -12
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -40
