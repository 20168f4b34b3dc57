/* hwloc is only included, never used, by the codelet bodies */
