cd /root/repo
timeout 2700 python -m pytest tests -q -m gpu 2>&1 | grep -v "^$" | grep -v "Warning\|warnings.warn\|^  " | tail -70
