# shim
