set -e
python profiles/probes/iic_micro.py --out /tmp/ref.npz
CY_IIC_TFORM=1 python profiles/probes/iic_micro.py --ref /tmp/ref.npz
python profiles/probes/iic_micro.py --K 9 --H 100 --W 52 --out /tmp/ref9.npz
CY_IIC_TFORM=1 python profiles/probes/iic_micro.py --K 9 --H 100 --W 52 --ref /tmp/ref9.npz
python profiles/probes/iic_micro.py --K 5 --H 37 --W 44 --B 3 --out /tmp/ref5.npz
CY_IIC_TFORM=1 python profiles/probes/iic_micro.py --K 5 --H 37 --W 44 --B 3 --ref /tmp/ref5.npz
