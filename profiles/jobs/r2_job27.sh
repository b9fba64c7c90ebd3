cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300
for w in td7_humanoid_pop256 td7_ant_pop256; do timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1; done
