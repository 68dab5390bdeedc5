python tools/agg_cluster_probe.py --batch 2048
