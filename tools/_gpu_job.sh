./tools/micro/gather_lat
