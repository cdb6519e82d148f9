#!/usr/bin/env bash
python scripts/host_profile.py c2p 2>&1 | cut -c1-170 | head -70
