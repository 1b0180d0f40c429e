#pragma once
/* intentionally empty: printing helpers are not needed by the checker build */
