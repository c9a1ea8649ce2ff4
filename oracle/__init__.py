"""ORACLE — test infrastructure only.

CPU restatement of the reference's hot path (CVMI-Lab/IST-Net, /root/reference):
  * pointops_ref.c / pointops.py : the nine `pointnet2._ext` operators (bindings.cpp:11-24)
  * istnet_port.py               : plain-PyTorch functional restatement of IST_Net / PoseNetGT
  * ref_harness.py               : imports the UNMODIFIED reference Python (only where /root/reference exists)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package. Nothing under istnet_b200/ does.
"""
