AMGB200_CTA_G=4 AMGB200_CTA_D=2 python tools/quick_time.py aniso3d 64 2>&1 | grep -E "^solve|^L[45]" | cut -c1-120
AMGB200_CTA_G=2 AMGB200_CTA_D=2 python tools/quick_time.py aniso3d 64 2>&1 | grep -E "^solve|^L[45]" | cut -c1-120
AMGB200_CTA_G=1 AMGB200_CTA_D=2 python tools/quick_time.py aniso3d 64 2>&1 | grep -E "^solve|^L[45]" | cut -c1-120
AMGB200_CTA_G=2 AMGB200_CTA_D=2 python tools/quick_time.py p2d 256 2>&1 | grep -E "^solve" | cut -c1-120
