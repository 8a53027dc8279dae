#pragma once
namespace ba { struct LocalDev { int n_windows; }; }
