cd /root/repo
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -60
