cd /root/repo
python tools/stress_parity.py 300 7 2>&1 | tail -2
VSL_NVCC_EXTRA=-DVSL_EXACT_RCP python -m unsupervised_pose_estimation_b200.build --force > /dev/null && python tools/stress_parity.py 300 7 2>&1 | tail -2
