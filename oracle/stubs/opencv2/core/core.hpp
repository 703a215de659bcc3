/* intentionally empty: unused include of the reference (oracle build only) */
