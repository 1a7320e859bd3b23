cd "$(dirname "$0")/.."; python tools/prof_track.py 296 500; python tools/prof_track.py 8 500
