"""Test infrastructure: CPU restatement of the reference hot path + loaders for the compiled reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (lidardetection_b200/) never does.
"""
