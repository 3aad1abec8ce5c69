# shim: plotting is never reached on the update path
